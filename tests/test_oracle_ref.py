"""Validate the oracle's restatement against the REFERENCE's own header-only numerics, compiled in place from
/root/reference into oracle/_ref/libhalmd_ref.so (oracle/Makefile, oracle/ref_shim.cpp).  Skipped when that
library was not built."""
import ctypes as C

import numpy as np

import support


def _ds_cases(rng, n):
    a = rng.uniform(-1, 1, n) * 10.0 ** rng.integers(-3, 4, n)
    hi = a.astype(np.float32)
    lo = (a - hi.astype(np.float64)).astype(np.float32)
    return np.stack([hi, lo], axis=1)


def test_dsfun_bitwise(oracle, ref):
    """halmd/numeric/mp/dsfun.hpp:142-207 — the oracle's dsadd/dssub/dsmul equal the reference's bit for bit
    (both compiled without FMA contraction)."""
    rng = np.random.default_rng(42)
    A, B = _ds_cases(rng, 20000), _ds_cases(rng, 20000)
    out_o = np.zeros(2, dtype=np.float32)
    out_r = np.zeros(2, dtype=np.float32)
    for name_o, name_r in (("oracle_ds_add", "ref_dsadd"), ("oracle_ds_sub", "ref_dssub"), ("oracle_ds_mul", "ref_dsmul")):
        fo, fr = getattr(oracle, name_o), getattr(ref, name_r)
        for a, b in zip(A, B):
            a = np.ascontiguousarray(a)
            b = np.ascontiguousarray(b)
            fo(support.ptr(a), support.ptr(b), support.ptr(out_o))
            fr(support.ptr(a), support.ptr(b), support.ptr(out_r))
            assert out_o.tobytes() == out_r.tobytes(), (name_o, a, b, out_o, out_r)


def test_dsfun_accuracy(oracle):
    """test/unit/numeric/mp/dsfun.cpp tolerances: add/sub 1e-15 (relative to operands), mul 2e-14."""
    rng = np.random.default_rng(7)
    A, B = _ds_cases(rng, 5000), _ds_cases(rng, 5000)
    out = np.zeros(2, dtype=np.float32)
    for a, b in zip(A, B):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        da, db = float(a[0]) + float(a[1]), float(b[0]) + float(b[1])
        oracle.oracle_ds_add(support.ptr(a), support.ptr(b), support.ptr(out))
        assert abs(float(out[0]) + float(out[1]) - (da + db)) <= 2e-14 * max(abs(da), abs(db))
        oracle.oracle_ds_mul(support.ptr(a), support.ptr(b), support.ptr(out))
        assert abs(float(out[0]) + float(out[1]) - da * db) <= 2e-13 * abs(da * db)


def test_hilbert_map_matches_reference_header(oracle, ref):
    """halmd/mdsim/sorts/hilbert_kernel.hpp:70-142,205-230 on unit-box coordinates, all depths."""
    rng = np.random.default_rng(3)
    pts = rng.uniform(-0.5, 0.5, (4000, 3)).astype(np.float32)
    edge = np.array([[-0.5, -0.5, -0.5], [0.5, 0.5, 0.5], [0, 0, 0], [-1e-9, 1e-9, 0.49999997], [0.25, -0.25, 0.125]],
                    dtype=np.float32)
    pts = np.concatenate([pts, edge])
    for depth in range(1, 11):
        for x, y, z in pts:
            assert oracle.oracle_hilbert_map_unit(float(x), float(y), float(z), depth) == \
                ref.ref_hilbert_map(float(x), float(y), float(z), depth)


def test_hilbert_map_is_bijection_on_lattice(oracle):
    """All 8^depth lattice cells receive distinct codes, and consecutive codes are face neighbours
    (up to the periodic wrap of the box)."""
    depth = 3
    n = 1 << depth
    codes = {}
    for ix in range(n):
        for iy in range(n):
            for iz in range(n):
                c = oracle.oracle_hilbert_map_unit((ix + 0.5) / n - 0.5, (iy + 0.5) / n - 0.5, (iz + 0.5) / n - 0.5, depth)
                assert c not in codes
                codes[c] = (ix, iy, iz)
    assert sorted(codes) == list(range(n ** 3))
    for c in range(n ** 3 - 1):
        a, b = np.array(codes[c]), np.array(codes[c + 1])
        d = np.abs(a - b)
        d = np.minimum(d, n - d)
        assert d.sum() == 1, (c, a, b)


def test_lj_functor_matches_reference_header(oracle, ref):
    """lennard_jones_kernel.hpp:43-56 (float and double instantiations)."""
    rng = np.random.default_rng(5)
    p = support.OLjParams()
    oracle.oracle_lj_init(C.byref(p), 1, 1, support.farr([0.7]), support.farr([1.3]), 0, support.farr([100.0]), C.c_float(0))
    for rr in rng.uniform(0.5, 9.0, 2000).astype(np.float32):
        fo, eo, fr, er = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        oracle.oracle_pair_f(C.byref(p), 0, C.c_float(rr), C.byref(fo), C.byref(eo))
        ref.ref_lj_compute_f(C.c_float(rr), C.c_float(p.sigma2[0]), C.c_float(p.epsilon[0]), C.byref(fr), C.byref(er))
        assert fo.value == fr.value and eo.value == er.value


def test_fcc_generator_matches_reference_lattice(ref):
    """support.fcc_positions uses the site rule of lattice_primitive.hpp:121-129."""
    out = np.zeros(3, dtype=np.float32)
    for n in range(32):
        ref.ref_fcc_site(n, 2, 2, 2, support.ptr(out))
        cell = n >> 2
        exp = [cell % 2 + (((n ^ (n >> 1)) & 1) * 2 + 1) / 4, (cell // 2) % 2 + ((n & 1) * 2 + 1) / 4,
               cell // 4 + ((n & 2) + 1) / 4]
        assert np.allclose(out, exp)


def test_welford_merge_matches_reference_accumulator(ref):
    """numeric/accumulator.hpp:56-132 — sequential accumulation + merge = what hx_msd_block's triples feed."""
    rng = np.random.default_rng(11)
    x = np.ascontiguousarray(rng.normal(3.0, 2.0, 1000))
    out = np.zeros(3)
    ref.ref_accumulate(support.ptr(x), 600, 400, support.ptr(out))
    assert out[0] == 1000
    assert abs(out[1] - x.mean()) < 1e-12
    assert abs(out[2] - x.var()) < 1e-12


BLOCKING_CASES = [
    # dt, max_lag, resolution, size, shift, separation, start step, steps
    (0.001, 0.12, 0.004, 5, 2, 3, 0, 240),
    (0.001, 1.0, 0.1, 10, 3, 10, 0, 2500),            # examples/liquid/lennard_jones.lua: size 10, shift = floor(sqrt(10))
    (0.001, 1.0, 0.1, 10, 3, 10, 700, 1730),          # scheme created in the middle of a run, odd run length
    (0.005, 3.0, 0.005, 4, 1, 4, 0, 700),             # shift 1: no odd levels
    (0.005, 3.0, 0.005, 4, 2, 1, 13, 700),            # separation below the interval
    (0.002, 0.9, 0.003, 3, 2, 50, 5, 900),            # resolution not a multiple of dt (rounded), large separation
    (0.001, 0.05, 0.01, 2, 1, 2, 0, 137),             # minimal block size
]


def test_blocking_schedule_against_reference_classes(ref):
    """observables/dynamics/blocking_scheme.cpp:34-200, observables/samples/blocking_scheme.hpp, correlation.hpp:208-224 compiled
    in place (oracle/ref_dynamics.cpp) and driven like sampler:run — the schedule restatement the GPU tests hold the
    product's blocking scheme to (support.blocking_reference) yields the same intervals and the same sequence of correlated
    sample pairs, with and without finalise()."""
    for dt, max_lag, res, size, shift, sep, start, steps in BLOCKING_CASES:
        for fin in (True, False):
            out = support.ref_blocking_schedule(ref, dt, max_lag, res, size, shift, sep, start, steps, finalise=fin)
            assert out is not None
            iv, ev, cnt, mean, time = out
            interval, events = support.blocking_reference(steps, dt, max_lag, res, size, shift, sep, start=start, finalise=fin)
            assert iv == interval
            mine = support.blocking_events(events)
            assert ev.shape == mine.shape and np.array_equal(ev, mine), (dt, max_lag, res, size, shift, sep, start, steps, fin)
            # accumulated per (level, lag): the count of pairs and the mean lag in steps
            exp = np.zeros_like(cnt)
            for lev, first, second in mine:
                exp[lev, (second - first) // interval[lev]] += 1
            assert np.array_equal(cnt, exp)
            nz = cnt > 0
            lag = np.array(interval)[:, None] * np.arange(size)[None, :]
            assert np.array_equal(mean[nz], lag[nz].astype(float))
            assert np.allclose(time, lag * dt, rtol=0, atol=0)


def test_blocking_scheme_constructor_errors(ref):
    """blocking_scheme.cpp:51-59,72-74,96-99: invalid arguments throw."""
    for args in ((0.001, 1.0, 0.1, 1, 1, 1), (0.001, 1.0, 0.1, 10, 0, 10), (0.001, 1.0, 0.1, 10, 10, 10),
                 (0.001, 1.0, -0.1, 10, 3, 10), (0.001, 0.01, 0.1, 10, 3, 10)):
        assert support.ref_blocking_schedule(ref, *args, 0, 10) is None


def test_isf_product_against_reference_functor(ref):
    """intermediate_scattering_function.hpp:90-113 on the shells of the reference's wavevector class
    (wavevector.cpp:40-123): the product's host-side hx_isf_accumulate gives the same accumulators bit for bit, also when
    several sample pairs are merged into the same output (correlation<>::compute over time); and the API's wavevector
    module equals the reference class (vectors, shells, discarded wavenumbers)."""
    from halmd_b200 import _lib as hx
    rng = np.random.default_rng(5)
    L = np.array([17.3, 21.1, 13.9])
    wavenumber = np.array([0.1, 0.75, 1.5, 2.0, 7.2, 12.5])        # (0.1 < 2 pi / L matches no lattice point: discarded)
    tol, max_count = 0.02, 7
    nq, nshell = C.c_uint(0), C.c_uint(0)
    fn = ref.ref_isf_accumulate
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_double, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint, C.c_double,
                   C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_void_p]
    assert fn(wavenumber.ctypes.data, len(wavenumber), L.ctypes.data, tol, max_count, None, None, 0, 1.0, C.byref(nq), C.byref(nshell), None) == 0
    nq, nshell = nq.value, nshell.value
    assert 0 < nshell < len(wavenumber)
    # the reference class: vectors, shells and kept wavenumbers
    vec = np.zeros((nq, 3))
    shells = np.zeros((nshell, 2), dtype=np.uint32)
    kept = np.zeros(nshell)
    fc = ref.ref_wavevector_class
    fc.restype = C.c_int
    fc.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_double, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
    assert fc(wavenumber.ctypes.data, len(wavenumber), L.ctypes.data, tol, max_count, vec.ctypes.data, nq, shells.ctypes.data, kept.ctypes.data) == nshell
    # the product's picker behind the C ABI with the caller-side drop of empty shells (api._wavevector)
    cap = len(wavenumber) * max_count
    pv = np.zeros((cap, 3))
    offs = np.zeros(len(wavenumber) + 1, dtype=np.uint32)
    hx.check(hx.lib.hx_wavevector_pick(wavenumber.ctypes.data_as(hx.D3), len(wavenumber), L.ctypes.data_as(hx.D3), tol, max_count,
                                       None, pv.ctypes.data_as(hx.D3), cap, offs.ctypes.data_as(hx.U3)))
    mine = [(int(offs[i]), int(offs[i + 1])) for i in range(len(wavenumber)) if offs[i + 1] > offs[i]]
    assert mine == [tuple(int(x) for x in s) for s in shells]
    assert pv[:nq].tobytes() == vec.tobytes()
    assert np.array_equal(kept, wavenumber[[i for i in range(len(wavenumber)) if offs[i + 1] > offs[i]]])
    shell_offsets = np.array([s[0] for s in mine] + [mine[-1][1]], dtype=np.uint32)
    for nsamples, norm in ((1, 4000.0), (7, 256000.0)):
        rho1 = rng.normal(size=(nsamples, nq, 2)) * 60
        rho2 = rho1 * 0.7 + rng.normal(size=(nsamples, nq, 2)) * 20
        out_ref = np.zeros((nshell, 3))
        assert fn(wavenumber.ctypes.data, len(wavenumber), L.ctypes.data, tol, max_count, rho1.ctypes.data, rho2.ctypes.data, nsamples,
                  norm, C.byref(C.c_uint()), C.byref(C.c_uint()), out_ref.ctypes.data) == 0
        acc = np.zeros((nshell, 3))
        for s in range(nsamples):
            a, b = np.ascontiguousarray(rho1[s]), np.ascontiguousarray(rho2[s])
            hx.check(hx.lib.hx_isf_accumulate(a.ctypes.data_as(hx.D3), b.ctypes.data_as(hx.D3), shell_offsets.ctypes.data_as(hx.U3),
                                              nshell, norm, acc.ctypes.data_as(hx.D3)))
        assert np.array_equal(acc[:, 0], out_ref[:, 0])
        assert acc[:, 1].tobytes() == out_ref[:, 1].tobytes()
        # variance = M2 / count (accumulator.hpp:150-158)
        assert np.allclose(acc[:, 2] / acc[:, 0], out_ref[:, 2], rtol=1e-15, atol=0)
