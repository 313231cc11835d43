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


def sorted_lists(s):
    """Device copy of the row-layout lists with every row sorted by particle index (experiment: gather locality)."""
    import numpy as np
    nb, n = s.neighbour, s.particle.nparticle
    raw = np.zeros(n * nb.stride, dtype=np.uint32)
    cnt = np.zeros(n, dtype=np.uint32)
    hx.check(hx.lib.hxm_neighbour_get(nb._h, raw.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
    rows = raw.reshape(n, nb.stride)
    k = np.arange(nb.stride)[None, :]
    rows = np.where(k < cnt[:, None], rows, 0xFFFFFFFF).astype(np.uint32)
    rows.sort(axis=1)
    return torch.from_numpy(rows.view(np.int32).copy()).cuda()


def kernel_lib():
    """HALMD_B200_KERNEL_LIB: a second build of the library used for the timed kernel call only (the system is set up
    and melted with the regular library), so that diagnostic kernel variants see the same physical state."""
    path = os.environ.get("HALMD_B200_KERNEL_LIB")
    if not path:
        return hx.lib
    lib = C.CDLL(path)
    for name in ("hx_pair_force", "hx_pair_force_shells"):
        getattr(lib, name).restype = getattr(hx.lib, name).restype
        getattr(lib, name).argtypes = getattr(hx.lib, name).argtypes
    return lib


def time_force(s, lanes, flush, iters=20, lists=None, flags=0):
    lib = kernel_lib()
    p, nb, pot = s.particle, s.neighbour, s.potential
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
        hx.check(lib.hx_pair_force(C.byref(par), pos, pos, (lists.data_ptr() if lists is not None else hx.lib.hxm_neighbour_device_ptr(nb._h)), nb.size, nb.stride, nb.layout | flags | hx.FORCE_LANES(lanes), n,
                                   L, 1, 0.5, force.data_ptr(), None, None, p.array_size, C.c_void_p(stream.cuda_stream)))
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts), min(ts)


def time_shells(s, lanes, flush, iters=10, flags=0, guarded=True):
    """The shell-row kernel on the module's private rows, with the module's r0 and control block (the real displacement
    bound of the current state; the step counter it bumps is reset by the next chunk)."""
    lib = kernel_lib()
    p, nb, pot = s.particle, s.neighbour, s.potential
    n = p.nparticle
    stream = torch.cuda.current_stream()
    fl = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda") if flush else None
    pos = p.device_ptr("position")
    force = torch.empty(p.array_size * 4, dtype=torch.float32, device="cuda")
    sh = hx.lib.hxm_neighbour_named_ptr(nb._h, b"shell_rows")
    if not sh:
        raise SystemExit("private shell rows are not in use")
    sh = C.cast(sh, C.POINTER(hx.ShellRows))
    r0 = hx.lib.hxm_neighbour_named_ptr(nb._h, b"r0") if guarded else None
    ctl = hx.lib.hxm_neighbour_named_ptr(nb._h, b"ctl") if guarded else None
    L = hx.f3(s.box.length)
    par = pot.params
    ts = []
    for i in range(iters + 3):
        if flush:
            fl.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        hx.check(lib.hx_pair_force_shells(C.byref(par), pos, pos, sh, r0, n, L, flags | hx.FORCE_LANES(lanes), 1, force.data_ptr(),
                                          ctl, 0.25, C.c_void_p(stream.cuda_stream)))
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts), min(ts), force


def shells_mode(a, cfg):
    """Row kernel (full walk of the public rows) vs shell kernel (prefix walk of the private rows) at several points of a
    rebuild interval of the melted system, L2 flushed and warm."""
    import numpy as np
    args = argparse.Namespace(layout="row", melt_steps=2000, lanes=4)
    s = bench.build_system(args, cfg, seed=1)
    s.run(300)
    nb = s.neighbour
    lane_set = (2, 4, 8) if a.lanes_sweep else (4,)
    red = hx.FORCE_POSITIONS_REDUCED
    # walk to the first step after a list update
    u0 = hx.lib.hxm_neighbour_updates(nb._h)
    while hx.lib.hxm_neighbour_updates(nb._h) == u0:
        s.run(1)
    since = 0
    while since <= 48:
        ctl = (C.c_uint32 * 4)()
        hx.check(hx.lib.hxm_neighbour_ctl_read(nb._h, ctl))
        d_max = float(np.sqrt(np.array([ctl[0]], dtype=np.uint32).view(np.float32)[0]))
        row = {"steps_since_update": since, "d_max": round(d_max, 4)}
        for flush in (True, False):
            tag = "flushed" if flush else "warm"
            avg, best = time_force(s, 4, flush, iters=10, flags=red)
            row[f"row2_{tag}_us"] = round(avg, 2)
            for lanes in lane_set:
                avg, best, f = time_shells(s, lanes, flush, flags=red)
                row[f"shell_l{lanes}_{tag}_us"] = round(avg, 2)
            avg, best, f = time_shells(s, 4, flush, flags=red, guarded=False)
            row[f"shell_fullrow_{tag}_us"] = round(avg, 2)
        print(json.dumps(row), flush=True)
        u = hx.lib.hxm_neighbour_updates(nb._h)
        s.run(6)
        since = since + 6 if hx.lib.hxm_neighbour_updates(nb._h) == u else 0
        if since == 0:
            break


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--sorted", action="store_true", help="also time the row kernel on index-sorted rows")
    ap.add_argument("--quick", action="store_true", help="row layout, packed kernel only: lanes x reduced-positions flag")
    ap.add_argument("--shells", action="store_true", help="shell-row kernel vs row kernel along a rebuild interval")
    ap.add_argument("--lanes-sweep", action="store_true")
    a = ap.parse_args()
    cfg = bench.CONFIGS[a.config]
    out = []
    if a.shells:
        return shells_mode(a, cfg)
    if a.quick:
        args = argparse.Namespace(layout="row", melt_steps=2000)
        s = bench.build_system(args, cfg, seed=1)
        s.run(200)
        s.neighbour.check()
        for lanes in (2, 4, 8):
            for flags in (0, hx.FORCE_POSITIONS_REDUCED):
                avg, best = time_force(s, lanes, True, flags=flags)
                print(json.dumps(dict(lib=os.path.basename(os.environ.get('HALMD_B200_KERNEL_LIB') or hx.LIB_PATH), lanes=lanes, reduced=int(flags != 0), avg_us=round(avg, 2),
                                      min_us=round(best, 2))), flush=True)
        if a.sorted:
            srt = sorted_lists(s)
            for lanes in (2, 4, 8):
                avg, best = time_force(s, lanes, True, lists=srt, flags=hx.FORCE_POSITIONS_REDUCED)
                print(json.dumps(dict(sorted_rows=True, lanes=lanes, avg_us=round(avg, 2), min_us=round(best, 2))), flush=True)
        return
    for layout in ("row", "column"):
        args = argparse.Namespace(layout=layout, melt_steps=2000)
        s = bench.build_system(args, cfg, seed=1)
        s.run(200)
        s.neighbour.check()
        nbar = s.neighbour.mean_length()
        for lanes in ((2, 4, 8) if layout == "row" else (8,)):
            for packed in ((1, 0) if layout == "row" else ()):
                avg, best = time_force(s, lanes, True, flags=0 if packed else hx.FORCE_SCALAR)
                print(json.dumps(dict(layout=layout, lanes=lanes, packed=packed, avg_us=avg, min_us=best)), flush=True)
            for flush in (True, False):
                avg, best = time_force(s, lanes, flush)
                out.append(dict(layout=layout, lanes=lanes, flush=flush, avg_us=avg, min_us=best, nbar=nbar,
                                gbs=cfg["n"] * (32 + 4 * nbar) / avg / 1e3))
                print(json.dumps(out[-1]), flush=True)
            if layout == "row" and a.sorted:
                avg, best = time_force(s, lanes, True, lists=sorted_lists(s))
                print(json.dumps(dict(layout=layout, lanes=lanes, flush=True, sorted_rows=True, avg_us=avg, min_us=best)), flush=True)
        del s


if __name__ == "__main__":
    main()
