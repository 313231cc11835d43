#!/usr/bin/env python
"""Kernel micro-benchmark (development aid): pair-force kernel variants on the melted cfg2 fluid."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from halmd_b200 import _lib as hx  # noqa: E402


def time_force(s, lanes, flush, iters=20):
    lib = hx.lib
    p, nb, pot = s.particle, s.neighbour, s.potential
    hx.check(lib.hx_pair_force_set_lanes(lanes))
    n = p.nparticle
    stream = torch.cuda.current_stream()
    fl = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda") if flush else None
    pos = p.device_ptr("position")
    force = torch.empty(p.array_size * 4, dtype=torch.float32, device="cuda")
    L = hx.f3(s.box.length)
    par = pot.params
    ts = []
    for i in range(iters + 3):
        if flush:
            fl.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        hx.check(lib.hx_pair_force(C.byref(par), pos, pos, lib.hxm_neighbour_device_ptr(nb._h), nb.size, nb.stride, nb.layout, n,
                                   L, 1, 0.5, force.data_ptr(), None, None, p.array_size, C.c_void_p(stream.cuda_stream)))
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    a = ap.parse_args()
    cfg = bench.CONFIGS[a.config]
    out = []
    for layout in ("row", "column"):
        args = argparse.Namespace(layout=layout, melt_steps=2000)
        s = bench.build_system(args, cfg, seed=1)
        s.run(200)
        s.neighbour.check()
        nbar = s.neighbour.mean_length()
        for lanes in ((2, 4, 8) if layout == "row" else (8,)):
            for flush in (True, False):
                avg, best = time_force(s, lanes, flush)
                out.append(dict(layout=layout, lanes=lanes, flush=flush, avg_us=avg, min_us=best, nbar=nbar,
                                gbs=cfg["n"] * (32 + 4 * nbar) / avg / 1e3))
                print(json.dumps(out[-1]), flush=True)
        del s


if __name__ == "__main__":
    main()
