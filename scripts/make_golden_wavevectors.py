#!/usr/bin/env python
"""Generates tests/golden/wavevectors.json: wavevector shells picked by the REFERENCE's own algorithm
(halmd/algorithm/host/pick_lattice_points.hpp compiled in place into oracle/_ref/libhalmd_ref.so, called as
observables/utility/wavevector.cpp:40-123 calls it) for a handful of boxes, grids, tolerances and filters.
Run in the development container (needs /root/reference at build time of oracle/_ref); the fixture travels."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import support  # noqa: E402

CASES = [
    dict(name="cubic_lj_4000", L=[16.796, 16.796, 16.796], q=[0.5, 1.0, 2.0, 4.0, 7.2, 10.0], tol=0.01, max_count=7, filter=None),
    dict(name="semilog_ka", L=[21.54, 21.54, 21.54], q="semilog:15:10", tol=0.01, max_count=7, filter=None),
    dict(name="cuboid", L=[12.0, 15.0, 18.0], q=[0.6, 1.3, 2.1, 3.3, 6.0], tol=0.05, max_count=12, filter=None),
    dict(name="slab_parallel", L=[70.0, 30.0, 30.0], q=[0.21, 0.42, 0.84, 1.68, 3.36, 6.72], tol=0.02, max_count=20, filter=[0, 1, 1]),
    dict(name="slab_normal", L=[70.0, 30.0, 30.0], q=[0.09, 0.18, 0.9, 6.75], tol=0.01, max_count=5, filter=[1, 0, 0]),
    dict(name="tight", L=[10.0, 10.0, 10.0], q=[0.62831853, 3.3, 5.0], tol=1e-6, max_count=7, filter=None),
]


def grid(case):
    q = case["q"]
    if isinstance(q, str):
        _, stop, dec = q.split(":")
        out, h, x, i = [], 2 * np.pi / max(case["L"]), 2 * np.pi / max(case["L"]), 0
        while x < float(stop):
            out.append(x)
            x += h
            i += 1
            if int(dec) > 0 and i % int(dec) == 0:
                h *= 2
        return out
    return [float(v) for v in q]


def pick(fn, case):
    q = grid(case)
    nw = len(q)
    cap = nw * case["max_count"]
    vec = np.zeros((cap, 3))
    offs = (C.c_uint * (nw + 1))()
    filt = None if case["filter"] is None else (C.c_uint * 3)(*case["filter"])
    rc = fn((C.c_double * nw)(*q), nw, (C.c_double * 3)(*case["L"]), C.c_double(case["tol"]), case["max_count"], filt,
            vec.ctypes.data_as(C.c_void_p), cap, offs)
    return rc, q, vec[:offs[nw]], list(offs)


def main():
    ref = support.load_ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libhalmd_ref.so not built")
    ref.ref_wavevector_pick.restype = C.c_int
    out = []
    for case in CASES:
        rc, q, vec, offs = pick(ref.ref_wavevector_pick, case)
        assert rc == 0
        out.append(dict(case, wavenumbers=q, shell_offsets=offs, vectors=[[float.hex(float(x)) for x in v] for v in vec]))
        print(case["name"], len(vec), "vectors in", sum(1 for a, b in zip(offs, offs[1:]) if b > a), "of", len(q), "shells")
    with open(os.path.join(ROOT, "tests", "golden", "wavevectors.json"), "w") as f:
        json.dump({"source": "halmd/algorithm/host/pick_lattice_points.hpp:62-160 via oracle/ref_shim.cpp ref_wavevector_pick",
                   "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
