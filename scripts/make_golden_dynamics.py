#!/usr/bin/env python
"""Generates tests/golden/dynamics.json from the REFERENCE's own host-side dynamics classes compiled in place into
oracle/_ref/libhalmd_ref.so (oracle/ref_dynamics.cpp: mdsim/clock.cpp, observables/dynamics/blocking_scheme.cpp,
observables/samples/blocking_scheme.hpp, observables/dynamics/correlation.hpp,
observables/dynamics/intermediate_scattering_function.hpp, observables/utility/wavevector.cpp):
  * blocking scheme driven like sampler:run — sampling intervals and the sequence of correlated sample pairs
    (level, first step, later step), with finalise();
  * the ISF functor on seeded density-mode samples — (count, mean, variance) per wavevector shell after merging several
    sample pairs.
Run in the development container (needs /root/reference at build time of oracle/_ref); the fixture travels."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import support  # noqa: E402

BLOCKING = [
    # dt, max_lag, resolution, size, shift, separation, start step, steps
    (0.001, 0.12, 0.004, 5, 2, 3, 0, 240),
    (0.001, 1.0, 0.1, 10, 3, 10, 700, 1730),
    (0.005, 3.0, 0.005, 4, 1, 4, 0, 300),
    (0.002, 0.9, 0.003, 3, 2, 50, 5, 400),
]
ISF = dict(L=[17.3, 21.1, 13.9], wavenumber=[0.1, 0.75, 1.5, 2.0, 7.2, 12.5], tol=0.02, max_count=7, norm=4000.0, nsamples=5, seed=11)


def isf_samples(nq, nsamples, seed):
    rng = np.random.default_rng(seed)
    rho1 = rng.normal(size=(nsamples, nq, 2)) * 60
    rho2 = rho1 * 0.7 + rng.normal(size=(nsamples, nq, 2)) * 20
    return rho1, rho2


def main():
    ref = support.load_ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libhalmd_ref.so not built")
    out = {"blocking": [], "isf": None}
    for case in BLOCKING:
        dt, max_lag, res, size, shift, sep, start, steps = case
        iv, ev, cnt, _, _ = support.ref_blocking_schedule(ref, dt, max_lag, res, size, shift, sep, start, steps, finalise=True)
        out["blocking"].append({"args": list(case), "intervals": iv, "events": ev.reshape(-1).tolist(), "count": cnt.astype(int).tolist()})
    L, q = np.array(ISF["L"]), np.array(ISF["wavenumber"])
    fn = ref.ref_isf_accumulate
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_double, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint, C.c_double,
                   C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_void_p]
    nq, nshell = C.c_uint(0), C.c_uint(0)
    assert fn(q.ctypes.data, len(q), L.ctypes.data, ISF["tol"], ISF["max_count"], None, None, 0, 1.0, C.byref(nq), C.byref(nshell), None) == 0
    rho1, rho2 = isf_samples(nq.value, ISF["nsamples"], ISF["seed"])
    res = np.zeros((nshell.value, 3))
    assert fn(q.ctypes.data, len(q), L.ctypes.data, ISF["tol"], ISF["max_count"], rho1.ctypes.data, rho2.ctypes.data, ISF["nsamples"],
              ISF["norm"], C.byref(C.c_uint()), C.byref(C.c_uint()), res.ctypes.data) == 0
    vec = np.zeros((nq.value, 3))
    shells = np.zeros((nshell.value, 2), dtype=np.uint32)
    kept = np.zeros(nshell.value)
    fc = ref.ref_wavevector_class
    fc.restype = C.c_int
    fc.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_double, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
    assert fc(q.ctypes.data, len(q), L.ctypes.data, ISF["tol"], ISF["max_count"], vec.ctypes.data, nq.value, shells.ctypes.data,
              kept.ctypes.data) == nshell.value
    out["isf"] = dict(ISF, nq=nq.value, shells=shells.tolist(), kept_wavenumbers=kept.tolist(),
                      result_hex=[[float(x).hex() for x in row] for row in res])
    path = os.path.join(ROOT, "tests", "golden", "dynamics.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
