"""Properties of the oracle that the reference's own tests check on its implementation:
binning property, radix sort == stable sort, Verlet ideal-gas invariants, rand48 == drand48 recurrence,
plus brute-force cross-checks for the parts the reference leaves unpinned (neighbour sets, displacement)."""
import ctypes as C

import numpy as np
import pytest

import support

L1 = np.array([16.796, 16.796, 16.796])


def plan(oracle, n, L, rc, skin, occ):
    ncell = (C.c_uint * 3)()
    cl = (C.c_float * 3)()
    cs = C.c_uint()
    oracle.oracle_binning_plan(support.d3(L), C.c_size_t(n), C.c_float(rc), C.c_float(skin), C.c_double(occ), ncell, cl,
                               C.byref(cs))
    return list(ncell), np.array(list(cl), dtype=np.float32), cs.value


def test_sizing_table(oracle):
    """Sizes of the BASELINE configs by the reference's formulas (golden table from SURVEY §8)."""
    for row in support.golden("lj_potential.json")["sizing_table"]["rows"]:
        L = (row["n"] / row["rho"]) ** (1 / 3)
        ncell, cl, cs = plan(oracle, row["n"], [L] * 3, row["rc"], row["skin"], row["occ_bin"])
        assert ncell == [row["ncell"]] * 3, row["name"]
        assert cs == row["cell_size"], row["name"]
        size = oracle.oracle_nlist_size(row["rc"], row["skin"], row["n"], L ** 3, row["occ_list"], cs)
        assert size == row["list_size"], row["name"]
        assert oracle.oracle_array_size(row["n"]) == row["array_size"]
        assert oracle.oracle_hilbert_depth(support.d3([L] * 3)) == row["depth"]


def test_binning_property(oracle):
    """test/unit/mdsim/binning.cpp:105-197 — fcc lattice in a 2:5:3 box, sine-compressed; every particle sits
    in cell floor(r * ncell / L) and is binned exactly once."""
    unit = 8.0
    L = np.array([2, 5, 3]) * unit
    n = 4 * 4 * 10 * 6
    for s in (0.0, 0.25, 0.5, 0.75, 1.0):
        r = support.fcc_positions(n, L) + L / 2
        r = r - s * L / (2 * np.pi) * np.sin(2 * np.pi * r / L) * 0.95   # compression towards the box origin
        r -= L / 2
        r -= np.rint(r / L) * L
        p, v, A = support.pack_state(r)
        ncell, cl, cs = plan(oracle, n, L, 2.5, 0.5, 0.5)
        cs *= 4
        g_cell = np.zeros(int(np.prod(ncell)) * cs, dtype=np.uint32)
        cnt = np.zeros(int(np.prod(ncell)), dtype=np.uint32)
        ovf = oracle.oracle_binning_update(support.ptr(p), C.c_size_t(n), support.farr(cl), support.u3(ncell), cs,
                                           support.ptr(g_cell), support.ptr(cnt))
        assert ovf == 0
        ids = g_cell[g_cell != support.PH]
        assert np.array_equal(np.sort(ids), np.arange(n))
        cells = g_cell.reshape(-1, cs)
        # cell 0 starts at r = 0: index = floor(r / l) mod ncell (binning_kernel.cu:45-48, test/unit/mdsim/binning.cpp:146)
        expected = np.floor(p[:n, :3].astype(np.float64) / L * np.array(ncell)).astype(np.int64) % np.array(ncell)
        lin = (expected[:, 2] * ncell[1] + expected[:, 1]) * ncell[0] + expected[:, 0]
        where = np.full(n, -1)
        for c in range(cells.shape[0]):
            row = cells[c][cells[c] != support.PH]
            assert np.all(np.diff(row.astype(np.int64)) > 0)      # ascending index inside a cell
            where[row] = c
        # particles within float rounding of a cell face may legitimately differ
        frac = p[:n, :3].astype(np.float64) / L * np.array(ncell)
        interior = np.all(np.abs(frac - np.rint(frac)) > 1e-4, axis=1)
        assert np.array_equal(where[interior], lin[interior])


def test_binning_overflow_flag(oracle):
    rng = np.random.default_rng(0)
    n = 4000
    r = support.melt_like_positions(n, L1, rng)
    r[:40] = r[0] + rng.normal(scale=1e-3, size=(40, 3))       # 40 particles in one cell of size 32
    p, v, A = support.pack_state(r)
    ncell, cl, cs = plan(oracle, n, L1, 2.5, 0.5, 0.5)
    g_cell = np.zeros(int(np.prod(ncell)) * 32, dtype=np.uint32)
    cnt = np.zeros(int(np.prod(ncell)), dtype=np.uint32)
    assert oracle.oracle_binning_update(support.ptr(p), C.c_size_t(n), support.farr(cl), support.u3(ncell), 32,
                                        support.ptr(g_cell), support.ptr(cnt)) == 1


@pytest.mark.parametrize("n", [0, 1, 2, 255, 256, 257, 4097, 100000])
def test_sort_is_stable_sort(oracle, n):
    """test/unit/algorithm/radix_sort.cpp:99-230 — equals std::stable_sort, values follow their keys."""
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 21, n, dtype=np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    k, v = keys.copy(), vals.copy()
    oracle.oracle_sort_pairs_stable(support.ptr(k), support.ptr(v), C.c_size_t(n))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


def brute_force_sets(p, n, L, rr_cut_skin, ntype):
    """O(N^2) neighbour sets with the same float arithmetic (numpy float32, fma emulated in float64)."""
    pos = p[:n, :3]
    sp = p[:n, 3].view(np.uint32)
    Lf = L.astype(np.float32)
    sets = []
    for i in range(n):
        d = pos[i] - pos                                   # float32 subtraction
        q = np.rint(d / Lf)                                # IEEE float32 division + rint
        # d - q*L as one FMA: exact product/sum in float64, single rounding to float32
        d = (d.astype(np.float64) - q.astype(np.float64) * Lf.astype(np.float64)).astype(np.float32)
        x, y, z = d[:, 0], d[:, 1], d[:, 2]
        t = (y * y)                                        # float32 product
        t = (x.astype(np.float64) * x.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
        rr = (z.astype(np.float64) * z.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
        cut = rr_cut_skin[sp[i] * ntype + sp]
        m = rr <= cut
        m[i] = False
        sets.append(set(np.nonzero(m)[0].tolist()))
    return sets


def test_neighbour_sets_vs_brute_force(oracle):
    """from_binning is unpinned by the reference's tests: pin the oracle against an independent O(N^2) search
    using the identical float arithmetic (both layouts, binary mixture)."""
    rng = np.random.default_rng(1)
    n = 1500
    L = np.array([12.0, 13.0, 12.5])
    r = support.melt_like_positions(n, L, rng, jitter=0.2)
    species = (np.arange(n) >= int(0.8 * n)).astype(np.uint32)
    p, v, A = support.pack_state(r, species)
    rc = np.array([2.5, 2.0, 2.0, 2.2], dtype=np.float32)
    skin = np.float32(0.5)
    rr_cut_skin = ((rc + skin).astype(np.float64) ** 2).astype(np.float32)
    ncell, cl, cs = plan(oracle, n, L, float(rc.max()), 0.5, 0.5)
    assert min(ncell) >= 3
    g_cell = np.zeros(int(np.prod(ncell)) * cs, dtype=np.uint32)
    cnt = np.zeros(int(np.prod(ncell)), dtype=np.uint32)
    assert oracle.oracle_binning_update(support.ptr(p), C.c_size_t(n), support.farr(cl), support.u3(ncell), cs,
                                        support.ptr(g_cell), support.ptr(cnt)) == 0
    size = 200
    sets = brute_force_sets(p, n, L, rr_cut_skin, 2)
    for row_layout, stride in ((0, A), (1, 224)):
        lst = np.zeros(size * A if not row_layout else n * stride, dtype=np.uint32)
        count = np.zeros(n, dtype=np.uint32)
        ovf = oracle.oracle_nlist_build(support.ptr(g_cell), support.u3(ncell), cs, support.ptr(p), C.c_size_t(n),
                                        support.farr(rr_cut_skin), 2, support.f3(L), support.ptr(lst), size,
                                        C.c_size_t(stride), row_layout, support.ptr(count))
        assert ovf == 0
        for i in range(n):
            got = lst[i * stride:i * stride + count[i]] if row_layout else lst[i:i + count[i] * A:A]
            assert len(set(got.tolist())) == count[i]
            assert set(got.tolist()) == sets[i], i
    assert 60 < count.mean() < 130


def test_displacement_max(oracle):
    rng = np.random.default_rng(2)
    n = 3000
    r0 = support.melt_like_positions(n, L1, rng)
    dr = rng.normal(scale=0.05, size=(n, 3))
    r = r0 + dr
    r -= np.rint(r / L1) * L1                      # some particles cross the periodic boundary
    p0, _, A = support.pack_state(r0)
    p1, _, _ = support.pack_state(r)
    rr = oracle.oracle_displacement_max(support.ptr(p1), support.ptr(p0), C.c_size_t(n), support.f3(L1))
    expect = (dr ** 2).sum(axis=1).max()
    assert abs(rr - expect) < 1e-4 * expect


def test_rand48_matches_drand48_recurrence(oracle):
    """test/unit/random/gpu/rand48.cpp — leapfrogged stream == sequential X_{k+1} = a X_k + c mod 2^48."""
    rng = support.ORand48()
    blocks, threads = 4, 64
    T = blocks * threads
    state = np.zeros(3 * T, dtype=np.uint16)
    oracle.oracle_rand48_seed(C.byref(rng), blocks, threads, 42, support.ptr(state))
    n = 5 * T + 17
    out = np.zeros(n, dtype=np.float32)
    oracle.oracle_rand48_uniform(C.byref(rng), support.ptr(state), support.ptr(out), C.c_size_t(n))
    a, c, mask = 0x5DEECE66D, 0xB, (1 << 48) - 1
    x = 0x330E | ((42 & 0xFFFF) << 16) | (((42 >> 16) & 0xFFFF) << 32)
    for k in range(n):
        x = (a * x + c) & mask
        u = np.float32(np.float32((x >> 32) & 0xFFFF) / np.float32(65536.0)
                       + np.float32((x >> 16) & 0xFFFF) / np.float32(4294967296.0))
        assert out[k] == u, k
    # limb arithmetic of uint48.hpp equals the 64-bit shortcut
    r = np.zeros(3, dtype=np.uint16)
    st = np.array([0x330E, 0xABCD, 0x1234], dtype=np.uint16)
    av = np.array([a & 0xFFFF, (a >> 16) & 0xFFFF, (a >> 32) & 0xFFFF], dtype=np.uint16)
    cv = np.array([c, 0, 0], dtype=np.uint16)
    oracle.oracle_muladd48_limbs(support.ptr(av), support.ptr(st), support.ptr(cv), support.ptr(r))
    y = (a * (0x330E | (0xABCD << 16) | (0x1234 << 32)) + c) & mask
    assert [int(v) for v in r] == [y & 0xFFFF, (y >> 16) & 0xFFFF, (y >> 32) & 0xFFFF]


def test_verlet_ideal_gas(oracle):
    """test/unit/mdsim/integrators/verlet.cpp:107-140 — without forces E_kin and v_cm are conserved exactly and
    unfolded positions move ballistically (double-single accuracy)."""
    rng = np.random.default_rng(3)
    n = 1000
    L = np.array([10.0, 10.0, 10.0])
    r = rng.uniform(-5, 5, (n, 3))
    vel = rng.normal(size=(n, 3))
    p, v, A = support.pack_state(r, vel=vel)
    img = np.zeros((A, 4), dtype=np.float32)
    f = np.zeros((A, 4), dtype=np.float32)
    v0 = v.copy()
    dt = np.float32(0.001)
    steps = 1000
    for _ in range(steps):
        oracle.oracle_verlet_integrate(support.ptr(p), support.ptr(img), support.ptr(v), support.ptr(f), C.c_size_t(A),
                                       C.c_float(dt), support.f3(L))
        oracle.oracle_verlet_finalize(support.ptr(v), support.ptr(f), C.c_size_t(A), C.c_float(dt))
    assert np.array_equal(v, v0)
    unfolded = (p[:n, :3].astype(np.float64) + p[A:A + n, :3]) + img[:n, :3] * L
    v_ds = v0[:n, :3].astype(np.float64) + v0[A:A + n, :3]
    r_ds = r.astype(np.float32).astype(np.float64) + (r - r.astype(np.float32)).astype(np.float32)
    expect = r_ds + v_ds * float(dt) * steps
    assert np.abs(unfolded - expect).max() < 5e-9
    assert np.abs(p[:n, :3]).max() <= 5.0 + 1e-6


class OHooverChain(C.Structure):
    _fields_ = [("xi", C.c_double * 2), ("v_xi", C.c_double * 2), ("mass_xi", C.c_double * 2),
                ("temperature", C.c_double), ("en_kin_target_2", C.c_double), ("dof", C.c_double)]


def test_hoover_kernels_properties(oracle):
    """verlet_nvt_hoover_kernel.cu:37-125 restated: with scale = 1 the integrate step IS the plain Verlet step; a
    power-of-two scale is exact in both words; a general scale agrees with the product formed in double to
    double-single accuracy (2^-44)."""
    rng = np.random.default_rng(5)
    n = 2000
    L = np.array([12.0, 12.0, 12.0])
    p, v, A = support.pack_state(rng.uniform(-6, 6, (n, 3)), vel=rng.normal(size=(n, 3)))
    img = np.zeros((A, 4), dtype=np.float32)
    f = np.zeros((A, 4), dtype=np.float32)
    f[:n, :3] = rng.normal(scale=10, size=(n, 3))
    dt = C.c_float(0.002)
    p1, v1, i1 = p.copy(), v.copy(), img.copy()
    p2, v2, i2 = p.copy(), v.copy(), img.copy()
    oracle.oracle_hoover_integrate(support.ptr(p1), support.ptr(i1), support.ptr(v1), support.ptr(f), C.c_size_t(A), dt,
                                   C.c_double(1.0), support.f3(L))
    oracle.oracle_verlet_integrate(support.ptr(p2), support.ptr(i2), support.ptr(v2), support.ptr(f), C.c_size_t(A), dt,
                                   support.f3(L))
    assert p1.tobytes() == p2.tobytes() and v1.tobytes() == v2.tobytes() and i1.tobytes() == i2.tobytes()
    w = v.copy()
    oracle.oracle_hoover_rescale(support.ptr(w), C.c_size_t(A), C.c_double(0.5))
    assert np.array_equal(w[:n, :3], v[:n, :3] * np.float32(0.5)) and np.array_equal(w[A:A + n, :3], v[A:A + n, :3] * np.float32(0.5))
    assert np.array_equal(w[:n, 3], v[:n, 3])                                  # masses untouched
    scale = 0.9987654321012345
    w = v.copy()
    oracle.oracle_hoover_rescale(support.ptr(w), C.c_size_t(A), C.c_double(scale))
    got = w[:n, :3].astype(np.float64) + w[A:A + n, :3]
    want = (v[:n, :3].astype(np.float64) + v[A:A + n, :3]) * scale
    assert np.abs(got - want).max() <= 2.0 ** -43 * np.abs(want).max()
    assert np.all(np.abs(w[A:A + n, :3]) <= np.spacing(np.abs(w[:n, :3])) / 2 * (1 + 1e-6))      # normalised double-single


def test_hoover_chain_properties(oracle):
    """verlet_nvt_hoover.cpp:79-95, :185-222: chain masses after Martyna et al.; a system exactly at the target kinetic
    energy with the chain at rest is a fixed point (scale = 1); a hot system is damped, a cold one heated; the two
    half-step propagations of a step are mirror images, so stepping back with -dt restores the chain."""
    n, T, fres, dt = 1500, 1.0, 5.0, 0.002
    c = OHooverChain()
    oracle.oracle_hoover_chain_propagate.restype = C.c_double
    oracle.oracle_hoover_chain_energy.restype = C.c_double
    oracle.oracle_hoover_chain_init(C.byref(c), C.c_size_t(n), 3, C.c_double(T), C.c_double(fres))
    omega_sq = (2 * np.pi * fres) ** 2
    assert c.mass_xi[0] == pytest.approx(3 * n * T / omega_sq, rel=1e-15) and c.mass_xi[1] == pytest.approx(T / omega_sq, rel=1e-15)
    assert c.en_kin_target_2 == 3 * n * T
    s = oracle.oracle_hoover_chain_propagate(C.byref(c), C.c_double(3 * n * T), C.c_double(dt))
    # at the target the first variable stays at rest; only the second feels -T / m1 (its own thermal drive)
    assert s == 1.0 and c.v_xi[0] == 0.0 and c.xi[0] == 0.0 and c.v_xi[1] < 0
    oracle.oracle_hoover_chain_init(C.byref(c), C.c_size_t(n), 3, C.c_double(T), C.c_double(fres))
    hot = [oracle.oracle_hoover_chain_propagate(C.byref(c), C.c_double(3 * n * 3.0), C.c_double(dt)) for _ in range(20)]
    assert all(x < 1 for x in hot) and hot[-1] < hot[0]
    assert oracle.oracle_hoover_chain_energy(C.byref(c), C.c_size_t(n)) > 0
    oracle.oracle_hoover_chain_init(C.byref(c), C.c_size_t(n), 3, C.c_double(T), C.c_double(fres))
    cold = [oracle.oracle_hoover_chain_propagate(C.byref(c), C.c_double(3 * n * 0.3), C.c_double(dt)) for _ in range(20)]
    assert all(x > 1 for x in cold)
    # time reversal of one propagation at constant kinetic energy scaled consistently
    oracle.oracle_hoover_chain_init(C.byref(c), C.c_size_t(n), 3, C.c_double(T), C.c_double(fres))
    c.v_xi[0], c.v_xi[1], c.xi[0], c.xi[1] = 0.3, -0.2, 0.01, 0.02
    before = (c.xi[0], c.xi[1], c.v_xi[0], c.v_xi[1])
    ek = 3 * n * 1.7
    s = oracle.oracle_hoover_chain_propagate(C.byref(c), C.c_double(ek), C.c_double(dt))
    s_back = oracle.oracle_hoover_chain_propagate(C.byref(c), C.c_double(ek * s * s), C.c_double(-dt))
    assert s * s_back == pytest.approx(1.0, abs=1e-14)
    assert (c.xi[0], c.xi[1], c.v_xi[0], c.v_xi[1]) == pytest.approx(before, abs=1e-13)


def test_andersen_statistics(oracle):
    """test/unit/mdsim/integrators/verlet_nvt_andersen.cpp:103-174 (statistical): with rate*dt = 1 every warp
    collides; velocities are N(0, sqrt(T)) and the RNG state advances."""
    n = 8192
    A = support.array_size(n)
    v = np.zeros((2 * A, 4), dtype=np.float32)
    v[:A, 3] = 1
    f = np.zeros((A, 4), dtype=np.float32)
    rng = support.ORand48()
    state = np.zeros(3 * 32 * 256, dtype=np.uint16)
    oracle.oracle_rand48_seed(C.byref(rng), 32, 256, 1, support.ptr(state))
    s0 = state.copy()
    T = 1.5
    oracle.oracle_andersen_finalize(support.ptr(v), support.ptr(f), C.c_size_t(n), C.c_size_t(A), C.c_float(0.01),
                                    C.c_float(np.sqrt(T)), C.c_float(1.0), C.byref(rng), support.ptr(state))
    vel = v[:n, :3]
    assert not np.array_equal(s0, state)
    assert abs(vel.var() - T) < 0.05
    assert np.abs(vel.mean(axis=0)).max() < 0.05
    assert np.all(v[A:, :3] == 0)
    # no collisions at zero probability: pure half kick
    v2 = v.copy()
    f[:n, 0] = 2.0
    oracle.oracle_andersen_finalize(support.ptr(v2), support.ptr(f), C.c_size_t(n), C.c_size_t(A), C.c_float(0.01),
                                    C.c_float(np.sqrt(T)), C.c_float(0.0), C.byref(rng), support.ptr(state))
    assert np.allclose(v2[:n, 0] - v[:n, 0], 0.01, atol=2e-6)
