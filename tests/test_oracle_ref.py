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
