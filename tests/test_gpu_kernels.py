"""Parity of the CUDA path (through the C ABI, include/halmd_b200.h) with the CPU oracle on the same seeded
inputs.  Integer / index results must be bit-exact; floating-point results within the stated tolerances."""
import ctypes as C

import numpy as np
import pytest

import support

pytestmark = pytest.mark.gpu

EPS_F = float(np.finfo(np.float32).eps)


class Dev:
    """Small helper: numpy <-> torch device tensors and raw pointers."""

    def __init__(self, torch):
        self.torch = torch
        self.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def up(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()
        t._np_dtype, t._np_shape = a.dtype, a.shape
        return t

    def zeros(self, shape, dtype):
        a = np.zeros(shape, dtype=dtype)
        return self.up(a)

    def down(self, t):
        self.torch.cuda.synchronize()
        return t.cpu().numpy().view(t._np_dtype).reshape(t._np_shape).copy()

    @staticmethod
    def p(t, offset_bytes=0):
        return C.c_void_p(t.data_ptr() + offset_bytes)


@pytest.fixture(scope="module")
def dev(cuda):
    return Dev(cuda)


def lj_fluid(n, rho, seed, species_frac=None, jitter=0.12):
    L = np.array([(n / rho) ** (1 / 3)] * 3)
    rng = np.random.default_rng(seed)
    r = support.melt_like_positions(n, L, rng, jitter=jitter)
    species = None
    if species_frac is not None:
        species = (np.arange(n) >= int(species_frac * n)).astype(np.uint32)
        rng.shuffle(species)
    vel = rng.normal(scale=1.0, size=(n, 3))
    p, v, A = support.pack_state(r, species, vel)
    return L, p, v, A


def plan(hx, n, L, rc_max, skin, occ_bin, occ_list):
    ncell, cl, cs = (C.c_uint * 3)(), (C.c_float * 3)(), C.c_uint()
    hx.check(hx.lib.hx_binning_plan(hx.d3(L), n, rc_max, skin, occ_bin, ncell, cl, C.byref(cs)))
    size = C.c_uint()
    hx.check(hx.lib.hx_nlist_plan(rc_max, skin, n, float(np.prod(L)), occ_list, cs.value, C.byref(size)))
    return list(ncell), np.array(list(cl), dtype=np.float32), cs.value, size.value


def gpu_binning(hx, dev, p, n, cl, ncell, cs):
    d_pos = dev.up(p)
    ncells = int(np.prod(ncell))
    d_cell = dev.zeros(ncells * cs, np.uint32)
    d_cnt = dev.zeros(ncells, np.uint32)
    d_idx = dev.zeros(max(n, 1), np.uint32)
    d_flag = dev.zeros(1, np.int32)
    hx.check(hx.lib.hx_binning_update(dev.p(d_pos), n, hx.farr(cl), hx.u3(ncell), cs, dev.p(d_cell), dev.p(d_cnt),
                                      dev.p(d_idx), dev.p(d_flag), dev.stream), "hx_binning_update")
    return d_pos, d_cell, dev.down(d_cell), dev.down(d_cnt), dev.down(d_idx), int(dev.down(d_flag)[0])


def oracle_binning(oracle, p, n, cl, ncell, cs):
    ncells = int(np.prod(ncell))
    g_cell = np.zeros(ncells * cs, dtype=np.uint32)
    cnt = np.zeros(ncells, dtype=np.uint32)
    ovf = oracle.oracle_binning_update(support.ptr(p), C.c_size_t(n), support.farr(cl), support.u3(ncell), cs,
                                       support.ptr(g_cell), support.ptr(cnt))
    idx = np.zeros(max(n, 1), dtype=np.uint32)
    oracle.oracle_cell_index(support.ptr(p), C.c_size_t(n), support.farr(cl), support.u3(ncell), support.ptr(idx))
    return g_cell, cnt, idx, ovf


# ---- a2/a3 binning ------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,rho,seed", [(4000, 0.8442, 1), (256000, 0.8442, 2), (30000, 1.2, 3)])
def test_binning_bit_exact(hx, oracle, dev, n, rho, seed):
    L, p, v, A = lj_fluid(n, rho, seed)
    ncell, cl, cs, _ = plan(hx, n, L, 2.5, 0.5, 0.5, 0.4)
    _, _, g_cell, cnt, idx, flag = gpu_binning(hx, dev, p, n, cl, ncell, cs)
    o_cell, o_cnt, o_idx, o_flag = oracle_binning(oracle, p, n, cl, ncell, cs)
    assert flag == o_flag == 0
    assert np.array_equal(idx[:n], o_idx[:n])
    assert np.array_equal(cnt, o_cnt)
    assert np.array_equal(g_cell, o_cell)


def test_binning_edge_cases(hx, oracle, dev):
    # particles exactly on cell faces, box faces and at +-L/2; overflow; ghosts; non-cubic box
    L = np.array([12.0, 15.0, 18.0])
    n = 3000
    rng = np.random.default_rng(5)
    r = rng.uniform(-0.5, 0.5, (n, 3)) * L
    ncell, cl, cs, _ = plan(hx, n, L, 2.5, 0.5, 0.5, 0.4)
    faces = (rng.integers(-2, 3, (200, 3)) * cl.astype(np.float64))
    r[:200] = faces
    r[200] = L / 2
    r[201] = -L / 2
    r[202] = [0.0, -0.0, 0.0]
    p, v, A = support.pack_state(r)
    _, _, g_cell, cnt, idx, flag = gpu_binning(hx, dev, p, n, cl, ncell, cs)
    o_cell, o_cnt, o_idx, o_flag = oracle_binning(oracle, p, n, cl, ncell, cs)
    assert flag == o_flag
    assert np.array_equal(idx[:n], o_idx[:n]) and np.array_equal(g_cell, o_cell)
    # overcrowded cell raises the flag on both sides
    r[:100] = r[300] + rng.normal(scale=1e-3, size=(100, 3))
    p, v, A = support.pack_state(r)
    *_, flag = gpu_binning(hx, dev, p, n, cl, ncell, 32)
    *_, o_flag = oracle_binning(oracle, p, n, cl, ncell, 32)
    assert flag == o_flag == 1


# ---- a4 Hilbert keys, radix sort, a5 rearrange -----------------------------------------------------------

@pytest.mark.parametrize("n,rho", [(4000, 0.8442), (256000, 0.8442)])
def test_hilbert_keys_bit_exact(hx, oracle, dev, n, rho):
    L, p, v, A = lj_fluid(n, rho, 7)
    p[5, :3] = [-1e-9, 1e-9, L[2] / 2]           # rounding edge: r/L - floor(r/L) == 1
    depth = hx.lib.hx_hilbert_depth(hx.d3(L))
    d_pos = dev.up(p)
    d_keys = dev.zeros(A, np.uint32)
    hx.check(hx.lib.hx_hilbert_keys(dev.p(d_pos), A, hx.f3(L), depth, dev.p(d_keys), dev.stream))
    keys = dev.down(d_keys)
    o_keys = np.zeros(A, dtype=np.uint32)
    oracle.oracle_hilbert_keys(support.ptr(p), C.c_size_t(A), support.f3(L), depth, support.ptr(o_keys))
    assert np.array_equal(keys, o_keys)
    assert np.all(keys[n:] == support.PH)


@pytest.mark.parametrize("n,bits", [(1, 8), (2, 32), (255, 21), (2048, 21), (2049, 24), (100000, 30), (1000003, 21)])
def test_radix_sort_stable(hx, oracle, dev, n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << min(bits, 31), n, dtype=np.uint32)
    if bits == 32:
        keys |= np.uint32(0x80000000)
    keys[: n // 3] = keys[0]                       # many duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    d_k, d_v = dev.up(keys), dev.up(vals)
    sb = hx.lib.hx_sort_scratch_bytes(n)
    d_s = dev.zeros(sb, np.uint8)
    hx.check(hx.lib.hx_sort_pairs_stable_u32(dev.p(d_k), dev.p(d_v), n, bits, dev.p(d_s), sb, dev.stream))
    k, v = dev.down(d_k), dev.down(d_v)
    ok, ov = keys.copy(), vals.copy()
    oracle.oracle_sort_pairs_stable(support.ptr(ok), support.ptr(ov), C.c_size_t(n))
    assert np.array_equal(k, ok) and np.array_equal(v, ov)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(v, order.astype(np.uint32))


def test_rearrange_bit_exact(hx, oracle, dev):
    n = 10000
    L, p, v, A = lj_fluid(n, 0.8442, 9)
    rng = np.random.default_rng(9)
    img = np.zeros((A, 4), dtype=np.float32)
    img[:n, :3] = rng.integers(-3, 4, (n, 3))
    ids = np.arange(A, dtype=np.uint32)
    ids[:n] = rng.permutation(n)
    index = rng.permutation(n).astype(np.uint32)
    outs = [np.zeros_like(p), np.zeros_like(v), np.zeros_like(img), np.zeros_like(ids), np.zeros_like(ids)]
    oracle.oracle_state_permute(support.ptr(index), C.c_size_t(n), C.c_size_t(A), support.ptr(p), support.ptr(outs[0]),
                                support.ptr(v), support.ptr(outs[1]), support.ptr(img), support.ptr(outs[2]),
                                support.ptr(ids), support.ptr(outs[3]), support.ptr(outs[4]))
    d = [dev.up(x) for x in (index, p, v, img, ids)]
    o = [dev.zeros(x.shape, x.dtype) for x in outs]
    hx.check(hx.lib.hx_state_permute(dev.p(d[0]), n, A, dev.p(d[1]), dev.p(o[0]), dev.p(d[2]), dev.p(o[1]), dev.p(d[3]),
                                     dev.p(o[2]), dev.p(d[4]), dev.p(o[3]), dev.p(o[4]), dev.stream))
    for got, exp in zip(o, outs):
        assert dev.down(got).tobytes() == exp.tobytes()


# ---- a6 displacement ------------------------------------------------------------------------------------------

def test_displacement_bit_exact(hx, oracle, dev):
    n = 50000
    L, p0, v, A = lj_fluid(n, 0.8442, 11)
    rng = np.random.default_rng(11)
    p1 = p0.copy()
    p1[:n, :3] += rng.normal(scale=0.05, size=(n, 3)).astype(np.float32)
    p1[:n, :3] -= (np.rint(p1[:n, :3] / L.astype(np.float32)) * L.astype(np.float32)).astype(np.float32)
    d0, d1 = dev.up(p0), dev.up(p1)
    d_r0 = dev.zeros((n, 4), np.float32)
    hx.check(hx.lib.hx_displacement_zero(dev.p(d0), dev.p(d_r0), n, dev.stream))
    assert dev.down(d_r0).tobytes() == p0[:n].tobytes()
    d_bits = dev.zeros(4, np.uint32)
    hx.check(hx.lib.hx_displacement_max(dev.p(d1), dev.p(d_r0), n, hx.f3(L), dev.p(d_bits), 1, dev.stream))
    rr = dev.down(d_bits).view(np.float32)[0]
    o_rr = oracle.oracle_displacement_max(support.ptr(p1), support.ptr(p0), C.c_size_t(n), support.f3(L))
    assert rr == np.float32(o_rr)
    assert hx.lib.hx_displacement_exceeds(int(dev.down(d_bits)[0]), 0.5) == int(np.sqrt(np.float32(o_rr)) > np.float32(0.25))


# ---- a7 neighbour lists ----------------------------------------------------------------------------------------

def build_lists(hx, oracle, dev, n, rho, seed, rc, skin, occ_bin, occ_list, species_frac=None, layout=0, fill_all=1,
                jitter=0.12, reduce=False):
    L, p, v, A = lj_fluid(n, rho, seed, species_frac, jitter)
    if reduce:
        # positions as the integrator leaves them: in the primary cell [-L/2, L/2]
        Lf = L.astype(np.float32)
        p[:n, :3] -= (np.rint(p[:n, :3] / Lf) * Lf).astype(np.float32)
        assert np.all(np.abs(p[:n, :3]) <= 0.5 * Lf * (1 + 1e-6))
    nt = 1 if species_frac is None else 2
    rc = np.asarray(rc, dtype=np.float32).reshape(-1)
    ncell, cl, cs, size = plan(hx, n, L, float(rc.max()), skin, occ_bin, occ_list)
    rr_cut_skin = ((rc + np.float32(skin)).astype(np.float64) ** 2).astype(np.float32)
    d_pos, d_cell, g_cell, cnt, idx, flag = gpu_binning(hx, dev, p, n, cl, ncell, cs)
    assert flag == 0
    stride = A if layout == 0 else ((size + 32) // 32) * 32
    nl_len = size * stride if layout == 0 else n * stride
    d_nl = dev.up(np.full(nl_len, 0xDEADBEEF, dtype=np.uint32))        # poison: stale data must never be read
    d_count = dev.zeros(n, np.uint32)
    d_flag = dev.zeros(1, np.int32)
    hx.check(hx.lib.hx_nlist_build(dev.p(d_cell), hx.u3(ncell), cs, dev.p(d_pos), n, hx.farr(rr_cut_skin), nt, nt,
                                   hx.f3(L), dev.p(d_nl), size, stride, layout, fill_all, dev.p(d_count), dev.p(d_flag),
                                   dev.stream), "hx_nlist_build")
    o_nl = np.zeros(nl_len, dtype=np.uint32)
    o_count = np.zeros(n, dtype=np.uint32)
    o_flag = oracle.oracle_nlist_build(support.ptr(g_cell), support.u3(ncell), cs, support.ptr(p), C.c_size_t(n),
                                       support.farr(rr_cut_skin), nt, support.f3(L), support.ptr(o_nl), size,
                                       C.c_size_t(stride), layout, support.ptr(o_count))
    return dict(L=L, p=p, v=v, A=A, n=n, nt=nt, size=size, stride=stride, layout=layout, d_pos=d_pos, d_nl=d_nl,
                nl=dev.down(d_nl), count=dev.down(d_count), flag=int(dev.down(d_flag)[0]), o_nl=o_nl, o_count=o_count,
                o_flag=o_flag, ncell=ncell, cs=cs, cl=cl, g_cell=g_cell)


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("n,rho,rc,skin,ob,ol,frac", [
    (4000, 0.8442, [2.5], 0.5, 0.5, 0.4, None),
    (32000, 1.2, [2.5, 2.0, 2.0, 2.2], 0.3, 0.7, 0.7, 0.8),
])
def test_neighbour_lists_bit_exact(hx, oracle, dev, layout, n, rho, rc, skin, ob, ol, frac):
    """Same list ARRAY as the oracle (order included), hence identical neighbour sets."""
    s = build_lists(hx, oracle, dev, n, rho, 21, rc, skin, ob, ol, frac, layout=layout, fill_all=1)
    assert s["flag"] == s["o_flag"] == 0
    assert np.array_equal(s["count"], s["o_count"])
    assert np.array_equal(s["nl"], s["o_nl"])
    assert 40 < s["count"].mean() < 160


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("n,rho,rc,skin,ob,ol,frac", [
    (4000, 0.8442, [2.5], 0.5, 0.04, 0.4, None),
    (8000, 1.2, [2.5, 2.0, 2.0, 2.2], 0.3, 0.04, 0.7, 0.8),
])
def test_neighbour_lists_large_cells(hx, oracle, dev, layout, n, rho, rc, skin, ob, ol, frac):
    """Cells too large for the shared-memory staging (cell_size >= 512, as the doubling of binning.cpp:216-219 produces for
    strongly inhomogeneous systems) are served from global memory with the same list arrays."""
    s = build_lists(hx, oracle, dev, n, rho, 27, rc, skin, ob, ol, frac, layout=layout, fill_all=1)
    assert s["cs"] >= 512
    assert s["flag"] == s["o_flag"] == 0
    assert np.array_equal(s["count"], s["o_count"])
    assert np.array_equal(s["nl"], s["o_nl"])


def test_neighbour_lists_256k_sets(hx, oracle, dev):
    """BASELINE config 2 size, row layout without full fill: entries + terminator block."""
    s = build_lists(hx, oracle, dev, 256000, 0.8442, 22, [2.5], 0.5, 0.5, 0.4, layout=1, fill_all=0)
    assert s["flag"] == s["o_flag"] == 0
    assert np.array_equal(s["count"], s["o_count"])
    nl = s["nl"].reshape(s["n"], s["stride"])
    o_nl = s["o_nl"].reshape(s["n"], s["stride"])
    k = np.arange(s["stride"])[None, :]
    valid = k < s["count"][:, None]
    assert np.array_equal(nl[valid], o_nl[valid])
    term_end = np.minimum((s["count"] + 32) // 32 * 32, s["stride"])
    term = (k >= s["count"][:, None]) & (k < term_end[:, None])
    assert np.all(nl[term] == support.PH)


def test_neighbour_overflow_flag(hx, oracle, dev):
    """List reaching `size` entries raises the flag on both sides (from_binning_kernel.cu:263-266)."""
    L, p, v, A = lj_fluid(4000, 0.8442, 23)
    n = 4000
    ncell, cl, cs, size = plan(hx, n, L, 2.5, 0.5, 0.5, 0.4)
    size = 64
    rr = np.array([9.0], dtype=np.float32)
    d_pos, d_cell, g_cell, *_ = gpu_binning(hx, dev, p, n, cl, ncell, cs)
    d_nl = dev.zeros(size * A, np.uint32)
    d_flag = dev.zeros(1, np.int32)
    hx.check(hx.lib.hx_nlist_build(dev.p(d_cell), hx.u3(ncell), cs, dev.p(d_pos), n, hx.farr(rr), 1, 1, hx.f3(L),
                                   dev.p(d_nl), size, A, 0, 1, None, dev.p(d_flag), dev.stream))
    o_nl = np.zeros(size * A, dtype=np.uint32)
    o_flag = oracle.oracle_nlist_build(support.ptr(g_cell), support.u3(ncell), cs, support.ptr(p), C.c_size_t(n),
                                       support.farr(rr), 1, support.f3(L), support.ptr(o_nl), size, C.c_size_t(A), 0, None)
    assert int(dev.down(d_flag)[0]) == o_flag == 1
    assert np.array_equal(dev.down(d_nl), o_nl)       # truncated lists keep the first `size` entries


# ---- a8-a11 pair force -----------------------------------------------------------------------------------------

def run_force(hx, oracle, dev, s, trunc, aux, eps, sigma, cut, h=0.005, force_zero=1, lanes=8, flags=0):
    n, A, nt = s["n"], s["A"], s["nt"]
    par = hx.LjParams()
    hx.check(hx.lib.hx_potential_lj_init(C.byref(par), nt, nt, hx.farr(eps), hx.farr(sigma), trunc, hx.farr(cut), h))
    opar = support.OLjParams()
    oracle.oracle_lj_init(C.byref(opar), nt, nt, support.farr(eps), support.farr(sigma), trunc, support.farr(cut), C.c_float(h))
    rng = np.random.default_rng(99)
    f0 = np.zeros((A, 4), dtype=np.float32)
    e0 = np.zeros(A, dtype=np.float32)
    s0 = np.zeros(6 * A, dtype=np.float32)
    if not force_zero:
        f0[:n, :3] = rng.normal(size=(n, 3))
        e0[:n] = rng.normal(size=n)
        s0[:] = rng.normal(size=6 * A)
    d_f, d_e, d_s = dev.up(f0), dev.up(e0), dev.up(s0)
    hx.check(hx.lib.hx_pair_force(C.byref(par), dev.p(s["d_pos"]), dev.p(s["d_pos"]), dev.p(s["d_nl"]), s["size"],
                                  s["stride"], s["layout"] | flags | hx.FORCE_LANES(lanes), n, hx.f3(s["L"]), force_zero, 0.5,
                                  dev.p(d_f), dev.p(d_e) if aux else None, dev.p(d_s) if aux else None, A, dev.stream),
             "hx_pair_force")
    of, oe, os_ = f0.copy(), e0.copy(), s0.copy()
    oracle.oracle_pair_force(C.byref(opar), support.ptr(s["p"]), support.ptr(s["p"]), support.ptr(s["o_nl"]), s["size"],
                             C.c_size_t(s["stride"]), s["layout"], C.c_size_t(n), support.f3(s["L"]), force_zero,
                             C.c_float(0.5), support.ptr(of), support.ptr(oe) if aux else None,
                             support.ptr(os_) if aux else None, C.c_size_t(A))
    return dev.down(d_f), dev.down(d_e), dev.down(d_s), of, oe, os_


# The 1e-6 force bar of the north star is held against the REFERENCE's own kernels (tests/test_gpu_reference_kernels.py,
# tol=1e-6).  The CPU oracle divides and takes roots in IEEE arithmetic where the reference build (and, mirroring it,
# the product) uses MUFU.RCP / MUFU.SQRT: the two arithmetics differ by up to ~1.3e-6 per particle (2 ulp of
# sigma^2 / rr, amplified 7x by r^-14), hence the wider bound for oracle comparisons.
ORACLE_TOL = 3e-6


def check_force(f, of, n, tol=ORACLE_TOL):
    """Per-particle force agreement.  The relative error is taken against the larger of |f_i| and the rms force
    (forces on individual particles cancel to near zero in a liquid)."""
    d = np.linalg.norm(f[:n, :3].astype(np.float64) - of[:n, :3].astype(np.float64), axis=1)
    mag = np.linalg.norm(of[:n, :3].astype(np.float64), axis=1)
    rms = np.sqrt((mag ** 2).mean())
    rel = d / np.maximum(mag, rms)
    assert rel.max() < tol, (rel.max(), int(rel.argmax()))
    # and the bulk of particles agrees to float rounding of the result itself
    assert np.median(d / np.maximum(mag, 1e-30)) < 4 * EPS_F


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("trunc", [0, 1, 2, 3])
def test_pair_force_lj_single_species(hx, oracle, dev, layout, trunc):
    s = build_lists(hx, oracle, dev, 4000, 0.8442, 31, [2.5], 0.5, 0.5, 0.4, layout=layout, fill_all=0)
    f, e, st, of, oe, os_ = run_force(hx, oracle, dev, s, trunc, aux=True, eps=[1.0], sigma=[1.0], cut=[2.5])
    n, A = s["n"], s["A"]
    check_force(f, of, n)
    assert np.all(f[:n, 3] == 0)
    # en_pot / stress are plain float sums in the reference: tolerance = a few ulp of the sum of |terms|
    assert np.abs(e[:n] - oe[:n]).max() < 2e-5 * np.abs(oe[:n]).max()
    assert abs(e[:n].astype(np.float64).sum() - oe[:n].astype(np.float64).sum()) < 1e-6 * abs(oe[:n].astype(np.float64).sum())
    stv, osv = st.reshape(6, A)[:, :n], os_.reshape(6, A)[:, :n]
    assert np.abs(stv - osv).max() < 2e-5 * np.abs(osv).max()
    vir, ovir = stv[:3].astype(np.float64).sum(), osv[:3].astype(np.float64).sum()
    assert abs(vir - ovir) < 1e-6 * abs(ovir)


@pytest.mark.parametrize("layout,lanes", [(0, 8), (1, 2), (1, 4), (1, 8)])
def test_pair_force_binary_mixture(hx, oracle, dev, layout, lanes):
    """Kob-Andersen parameters, smooth_r4, skin 0.3 / occupancy 0.7 (benchmark settings)."""
    eps, sigma = [1.0, 1.5, 1.5, 0.5], [1.0, 0.8, 0.8, 0.88]
    rc = [2.5 * x for x in sigma]
    s = build_lists(hx, oracle, dev, 32000, 1.2, 33, rc, 0.3, 0.7, 0.7, species_frac=0.8, layout=layout, fill_all=0,
                    jitter=0.08)
    f, e, st, of, oe, os_ = run_force(hx, oracle, dev, s, 2, aux=False, eps=eps, sigma=sigma, cut=[2.5] * 4, lanes=lanes)
    check_force(f, of, s["n"])


@pytest.mark.parametrize("lanes", [2, 4, 8])
@pytest.mark.parametrize("trunc", [0, 1, 2, 3])
@pytest.mark.parametrize("eps,sigma", [(1.0, 1.0), (1.3, 0.9)])
def test_pair_force_packed_row_kernel(hx, oracle, dev, lanes, trunc, eps, sigma):
    """The production kernel (row layout, no aux, one species: packed f32x2 stream) in every truncation, lane count
    and parameter mode (epsilon = sigma = 1 selects the lennard_jones_simple arithmetic), and with the
    reduced-positions fast path, which must not change a single bit."""
    s = build_lists(hx, oracle, dev, 16000, 0.8442, 41 + lanes, [2.5 * sigma], 0.5, 0.5, 0.4, layout=1, fill_all=0, reduce=True)
    f, e, st, of, oe, os_ = run_force(hx, oracle, dev, s, trunc, aux=False, eps=[eps], sigma=[sigma], cut=[2.5], lanes=lanes)
    check_force(f, of, s["n"])
    f2, *_ = run_force(hx, oracle, dev, s, trunc, aux=False, eps=[eps], sigma=[sigma], cut=[2.5], lanes=lanes,
                       flags=hx.FORCE_POSITIONS_REDUCED)
    assert f2.tobytes() == f.tobytes()
    # compensated sums: the typical particle agrees with the oracle to float rounding of the result
    d = np.linalg.norm(f[:s["n"], :3].astype(np.float64) - of[:s["n"], :3].astype(np.float64), axis=1)
    mag = np.linalg.norm(of[:s["n"], :3].astype(np.float64), axis=1)
    assert np.mean(d / np.maximum(mag, 1e-30)) < 8 * EPS_F


def test_pair_force_accumulate_onto_existing(hx, oracle, dev):
    """force_zero = false adds to the previous force / aux arrays (pair_trunc_kernel.cuh:241-248)."""
    s = build_lists(hx, oracle, dev, 4000, 0.8442, 35, [2.5], 0.5, 0.5, 0.4, layout=1, fill_all=0)
    f, e, st, of, oe, os_ = run_force(hx, oracle, dev, s, 2, aux=True, eps=[1.0], sigma=[1.0], cut=[2.5], force_zero=0)
    check_force(f, of, s["n"])
    assert np.abs(e[:s["n"]] - oe[:s["n"]]).max() < 2e-5 * np.abs(oe[:s["n"]]).max()


def test_pair_force_256k(hx, oracle, dev):
    """BASELINE config 2 size; plus a size-independent property: Newton's third law, sum_i f_i = 0."""
    s = build_lists(hx, oracle, dev, 256000, 0.8442, 37, [2.5], 0.5, 0.5, 0.4, layout=1, fill_all=0)
    f, e, st, of, oe, os_ = run_force(hx, oracle, dev, s, 2, aux=False, eps=[1.0], sigma=[1.0], cut=[2.5])
    check_force(f, of, s["n"])
    tot = f[:s["n"], :3].astype(np.float64).sum(axis=0)
    scale = np.abs(f[:s["n"], :3]).astype(np.float64).sum()
    assert np.abs(tot).max() < 1e-6 * scale


def test_pair_potential_golden_chain(hx, oracle, dev):
    """The reference's GPU potential test (lennard_jones.cpp:392-436, smooth_r4.cpp:193-240): particles on a
    chain, each with exactly one neighbour; kernel output against the golden tables of the unit tests."""
    g = support.golden("lj_potential.json")
    eps, sigma = [1.0, 0.5, 0.5, 0.25], [1.0, 2.0, 2.0, 4.0]
    idx = {"aa": (0, 0), "ab": (0, 1), "bb": (1, 1)}
    L = np.array([100.0, 100.0, 100.0])
    for trunc_name, trunc in (("sharp", 0), ("shifted", 1), ("smooth_r4", 2), ("force_shifted", 3)):
        rows = []
        for pair, (ta, tb) in idx.items():
            for r, fval, en in g["lennard_jones"][trunc_name][pair]:
                rows.append((ta, tb, r, fval, en, 5.0 * g["pairs"][pair]["sigma"]))
        n = 2 * len(rows)
        pos = np.zeros((n, 3))
        species = np.zeros(n, dtype=np.uint32)
        nl = np.full((n, 32), support.PH, dtype=np.uint32)
        for k, (ta, tb, r, fval, en, rc) in enumerate(rows):
            # lists define the interactions, so all pairs may sit at the origin (keeps float rounding of r minimal)
            pos[2 * k] = [0.0, 0.0, 0.0]
            pos[2 * k + 1] = [r, 0.0, 0.0]
            species[2 * k], species[2 * k + 1] = ta, tb
            nl[2 * k, 0] = 2 * k + 1
        p, v, A = support.pack_state(pos, species)
        par = hx.LjParams()
        hx.check(hx.lib.hx_potential_lj_init(C.byref(par), 2, 2, hx.farr(eps), hx.farr(sigma), trunc, hx.farr([5.0] * 4), 0.005))
        d_pos, d_nl = dev.up(p), dev.up(nl)
        d_f, d_e, d_s = dev.zeros((A, 4), np.float32), dev.zeros(A, np.float32), dev.zeros(6 * A, np.float32)
        hx.check(hx.lib.hx_pair_force(C.byref(par), dev.p(d_pos), dev.p(d_pos), dev.p(d_nl), 31, 32, 1, n, hx.f3(L), 1, 1.0,
                                      dev.p(d_f), dev.p(d_e), dev.p(d_s), A, dev.stream))
        f, e = dev.down(d_f), dev.down(d_e)
        for k, (ta, tb, r, fval, en, rc) in enumerate(rows):
            dx = float(np.float32(p[2 * k, 0]) - np.float32(p[2 * k + 1, 0]))
            r_eff = abs(dx)                                   # positions are rounded to float
            if r >= rc:
                assert f[2 * k, 0] == 0 and e[2 * k] == 0
                continue
            # tolerance of the reference's GPU tests + the r^-14 amplification of the float position rounding
            tol = 10 * EPS_F + 14 * abs(r_eff - r) / r + 2e-6
            if trunc == 2:
                tol += 8 * EPS_F * (1 + rc / (rc - r))
            if trunc == 3:
                tol += 5e2 * 10 * EPS_F        # lennard_jones.cpp:266-271
            assert abs(f[2 * k, 0] - fval * dx) <= tol * abs(fval * dx) + 1e-30, (trunc_name, k, f[2 * k, 0], fval * dx)
            ref_en = abs(en) if abs(en) > 1e-3 * 4 * eps[2 * ta + tb] * (sigma[2 * ta + tb] / r) ** 12 else 4 * eps[2 * ta + tb] * (sigma[2 * ta + tb] / r) ** 12
            assert abs(e[2 * k] - en) <= tol * ref_en + 1e-30, (trunc_name, k, e[2 * k], en)


# ---- a12 velocity-Verlet ------------------------------------------------------------------------------------------

def test_hoover_kernels_bit_exact(hx, oracle, dev):
    """Nose-Hoover integrate(scale) / rescale(scale) against the oracle (verlet_nvt_hoover_kernel.cu:37-125): the
    double-single product v * scale with both low words live, then the Verlet half-step — hi and lo words, images."""
    n = 20000
    L, p, v, A = lj_fluid(n, 0.8442, 43)
    rng = np.random.default_rng(43)
    v[:n, :3] *= 30
    v[:n, 3] = rng.choice([1.0, 2.0, 0.5], n)
    v[A:A + n, :3] = (rng.normal(size=(n, 3)) * 1e-7).astype(np.float32)
    f = np.zeros((A, 4), dtype=np.float32)
    f[:n, :3] = rng.normal(scale=20, size=(n, 3))
    img = np.zeros((A, 4), dtype=np.float32)
    dt = 0.005
    d_p, d_v, d_f, d_i = dev.up(p), dev.up(v), dev.up(f), dev.up(img)
    op, ov, oi = p.copy(), v.copy(), img.copy()
    for scale in (0.9991234567891234, 1.0004321098765432, 1.0, 0.75):
        hx.check(hx.lib.hx_hoover_integrate(dev.p(d_p), dev.p(d_i), dev.p(d_v), dev.p(d_f), A, dt, scale, hx.f3(L), dev.stream))
        hx.check(hx.lib.hx_verlet_finalize(dev.p(d_v), dev.p(d_f), A, dt, dev.stream))
        hx.check(hx.lib.hx_hoover_rescale(dev.p(d_v), A, scale, dev.stream))
        oracle.oracle_hoover_integrate(support.ptr(op), support.ptr(oi), support.ptr(ov), support.ptr(f), C.c_size_t(A),
                                       C.c_float(dt), C.c_double(scale), support.f3(L))
        oracle.oracle_verlet_finalize(support.ptr(ov), support.ptr(f), C.c_size_t(A), C.c_float(dt))
        oracle.oracle_hoover_rescale(support.ptr(ov), C.c_size_t(A), C.c_double(scale))
    gp, gv, gi = dev.down(d_p), dev.down(d_v), dev.down(d_i)
    for name, got, exp in (("velocity", gv, ov), ("position", gp, op), ("image", gi, oi)):
        bad = np.argwhere((got[:n].view(np.uint32) != exp[:n].view(np.uint32)) & ~((got[:n] == 0) & (exp[:n] == 0)))
        assert bad.size == 0, (name, len(bad), bad[:4].tolist())
        bad = np.argwhere((got[A:A + n].view(np.uint32) != exp[A:A + n].view(np.uint32)) & ~((got[A:A + n] == 0) & (exp[A:A + n] == 0))) \
            if name != "image" else np.zeros(0)
        assert bad.size == 0, (name + " (low words)", len(bad))
    assert np.abs(gi[:n, :3]).max() >= 1


def test_verlet_bit_exact(hx, oracle, dev):
    """Double-single integrate/finalize: hi AND lo words equal the oracle's bit for bit, including periodic wrap
    and image bookkeeping; the fused kernel equals the separate calls."""
    n = 20000
    L, p, v, A = lj_fluid(n, 0.8442, 41)
    rng = np.random.default_rng(41)
    v[:n, :3] *= 30                                # fast particles: many boundary crossings within a few steps
    v[:n, 3] = rng.choice([1.0, 2.0, 0.5], n)      # masses
    f = np.zeros((A, 4), dtype=np.float32)
    f[:n, :3] = rng.normal(scale=20, size=(n, 3))
    img = np.zeros((A, 4), dtype=np.float32)
    dt = 0.005
    d_p, d_v, d_f, d_i = dev.up(p), dev.up(v), dev.up(f), dev.up(img)
    d_p2, d_v2, d_i2 = dev.up(p), dev.up(v), dev.up(img)
    op, ov, oi = p.copy(), v.copy(), img.copy()
    for step in range(5):
        hx.check(hx.lib.hx_verlet_integrate(dev.p(d_p), dev.p(d_i), dev.p(d_v), dev.p(d_f), A, dt, hx.f3(L), dev.stream))
        hx.check(hx.lib.hx_verlet_finalize(dev.p(d_v), dev.p(d_f), A, dt, dev.stream))
        # fused path: finalize of the previous step + integrate of this one
        hx.check(hx.lib.hx_verlet_finalize_integrate(dev.p(d_p2), dev.p(d_i2), dev.p(d_v2), dev.p(d_f), n, A, dt, hx.f3(L),
                                                     1 if step > 0 else 0, None, None, dev.stream))
        oracle.oracle_verlet_integrate(support.ptr(op), support.ptr(oi), support.ptr(ov), support.ptr(f), C.c_size_t(A),
                                       C.c_float(dt), support.f3(L))
        oracle.oracle_verlet_finalize(support.ptr(ov), support.ptr(f), C.c_size_t(A), C.c_float(dt))
    hx.check(hx.lib.hx_verlet_finalize(dev.p(d_v2), dev.p(d_f), A, dt, dev.stream))
    gp, gv, gi = dev.down(d_p), dev.down(d_v), dev.down(d_i)

    def same(name, got, exp):
        # bit patterns (species tags of ghost slots are NaN as floats); +0 and -0 count as equal
        bad = np.argwhere((got.view(np.uint32) != exp.view(np.uint32)) & ~((got == 0) & (exp == 0)))
        assert bad.size == 0, (name, len(bad), bad[:4].tolist(), [(got[tuple(b)], exp[tuple(b)]) for b in bad[:4]])

    same("velocity", gv, ov)
    same("position", gp, op)
    same("image", gi, oi)
    assert np.abs(gi[:n, :3]).max() >= 1            # the wrap path was exercised
    same("fused position", dev.down(d_p2), op)
    same("fused velocity", dev.down(d_v2), ov)
    same("fused image", dev.down(d_i2), oi)


def test_fused_integrate_displacement_and_guard(hx, oracle, dev):
    """Fused max-displacement equals the stand-alone reduction; a halted control block freezes the state."""
    n = 30000
    L, p, v, A = lj_fluid(n, 0.8442, 43)
    f = np.zeros((A, 4), dtype=np.float32)
    img = np.zeros((A, 4), dtype=np.float32)
    d_p, d_v, d_f, d_i = dev.up(p), dev.up(v), dev.up(f), dev.up(img)
    d_r0 = dev.up(p[:n].copy())
    d_ctl = dev.zeros(4, np.uint32)
    dt = 0.01
    hx.check(hx.lib.hx_verlet_finalize_integrate(dev.p(d_p), dev.p(d_i), dev.p(d_v), dev.p(d_f), n, A, dt, hx.f3(L), 0,
                                                 dev.p(d_r0), dev.p(d_ctl), dev.stream))
    p1 = dev.down(d_p)
    ctl = dev.down(d_ctl)
    o_rr = oracle.oracle_displacement_max(support.ptr(p1), support.ptr(p), C.c_size_t(n), support.f3(L))
    assert ctl[0] == np.float32(o_rr).view(np.uint32)
    # halted: nothing moves
    ctl[1] = 1
    d_ctl = dev.up(ctl)
    hx.check(hx.lib.hx_verlet_finalize_integrate(dev.p(d_p), dev.p(d_i), dev.p(d_v), dev.p(d_f), n, A, dt, hx.f3(L), 1,
                                                 dev.p(d_r0), dev.p(d_ctl), dev.stream))
    assert dev.down(d_p).tobytes() == p1.tobytes()


# ---- a13 rand48 + Andersen --------------------------------------------------------------------------------------

def test_rand48_bit_exact(hx, oracle, dev):
    for seed in (0, 1, 42, 0xDEADBEEF):
        rng, orng = hx.Rand48(), support.ORand48()
        d_state = dev.zeros(3 * 32 * 256, np.uint16)
        hx.check(hx.lib.hx_rand48_seed(C.byref(rng), 32, 256, seed, dev.p(d_state), dev.stream))
        o_state = np.zeros(3 * 32 * 256, dtype=np.uint16)
        oracle.oracle_rand48_seed(C.byref(orng), 32, 256, seed, support.ptr(o_state))
        assert np.array_equal(dev.down(d_state), o_state)
        assert list(rng.a) == list(orng.a) and list(rng.c) == list(orng.c)
        n = 100003
        d_out = dev.zeros(n, np.float32)
        hx.check(hx.lib.hx_rand48_uniform(C.byref(rng), dev.p(d_state), dev.p(d_out), n, dev.stream))
        o_out = np.zeros(n, dtype=np.float32)
        oracle.oracle_rand48_uniform(C.byref(orng), support.ptr(o_state), support.ptr(o_out), C.c_size_t(n))
        assert np.array_equal(dev.down(d_out), o_out)
        assert np.array_equal(dev.down(d_state), o_state)


@pytest.mark.parametrize("n", [4000, 100000])
def test_andersen_finalize(hx, oracle, dev, n):
    """Same collision pattern (which warps redraw) and RNG stream as the oracle; redrawn velocities agree to float
    rounding of log/sqrt; deterministic particles are bit-exact."""
    L, p, v, A = lj_fluid(n, 0.8442, 51)
    rng_np = np.random.default_rng(51)
    f = np.zeros((A, 4), dtype=np.float32)
    f[:n, :3] = rng_np.normal(scale=5, size=(n, 3))
    rng, orng = hx.Rand48(), support.ORand48()
    d_state = dev.zeros(3 * 32 * 256, np.uint16)
    hx.check(hx.lib.hx_rand48_seed(C.byref(rng), 32, 256, 7, dev.p(d_state), dev.stream))
    o_state = np.zeros(3 * 32 * 256, dtype=np.uint16)
    oracle.oracle_rand48_seed(C.byref(orng), 32, 256, 7, support.ptr(o_state))
    d_v, d_f = dev.up(v), dev.up(f)
    ov = v.copy()
    dt, T, prob = 0.005, 0.7, 2.0 * 0.005 * 10
    sqT = float(np.sqrt(np.float32(T)))

    def both():
        hx.check(hx.lib.hx_andersen_finalize(dev.p(d_v), dev.p(d_f), n, A, dt, sqT, prob, C.byref(rng), dev.p(d_state),
                                             dev.stream))
        oracle.oracle_andersen_finalize(support.ptr(ov), support.ptr(f), C.c_size_t(n), C.c_size_t(A), C.c_float(dt),
                                        C.c_float(sqT), C.c_float(prob), C.byref(orng), support.ptr(o_state))

    both()
    gv = dev.down(d_v)
    assert np.array_equal(dev.down(d_state), o_state)                 # identical RNG consumption
    redrawn = np.all(ov[A:A + n, :3] == 0, axis=1)                    # low words are cleared by a collision
    assert 0.02 < redrawn.mean() < 0.5
    det = ~redrawn
    assert np.array_equal(gv[:n][det], ov[:n][det]) and np.array_equal(gv[A:A + n][det], ov[A:A + n][det])
    assert np.allclose(gv[:n][redrawn, :3], ov[:n][redrawn, :3], rtol=2e-6, atol=1e-7)
    assert np.all(gv[A:A + n][redrawn, :3] == 0)
    assert abs(gv[:n][redrawn, :3].var() - T) < 0.05
    # further rounds: the streams stay in lock-step (redrawn values differ by rounding of log/sqrt only)
    for _ in range(3):
        both()
    assert np.array_equal(dev.down(d_state), o_state)
    gv = dev.down(d_v)
    assert np.allclose(gv[:n, :3], ov[:n, :3], rtol=4e-6, atol=1e-6)


@pytest.mark.parametrize("n", [1000, 8192 * 3 + 77, 300000])
def test_andersen_two_pass_equals_one_pass(hx, dev, n):
    """hx_andersen_finalize2 (plan on the generator grid + one thread per particle) reproduces the reference-shaped
    one-pass kernel bit for bit over several rounds: velocities (high and low words) and generator state."""
    L, p, v, A = lj_fluid(n, 0.8442, 57)
    rng_np = np.random.default_rng(57)
    v[:n, 3] = rng_np.choice(np.array([1.0, 1.7, 0.6], dtype=np.float32), n)
    f = np.zeros((A, 4), dtype=np.float32)
    f[:n, :3] = rng_np.normal(scale=5, size=(n, 3))
    rng = hx.Rand48()
    d_state = dev.zeros(3 * 32 * 256, np.uint16)
    hx.check(hx.lib.hx_rand48_seed(C.byref(rng), 32, 256, 13, dev.p(d_state), dev.stream))
    d_state2 = dev.up(dev.down(d_state))
    d_v, d_v2, d_f = dev.up(v), dev.up(v), dev.up(f)
    sb = hx.lib.hx_andersen_scratch_bytes(n)
    d_scr = dev.up(np.full(sb, 0xAB, dtype=np.uint8))                 # poison: nothing stale may be read
    dt, sqT, prob = 0.005, float(np.sqrt(np.float32(0.7))), 0.1
    for rnd in range(4):
        hx.check(hx.lib.hx_andersen_finalize(dev.p(d_v), dev.p(d_f), n, A, dt, sqT, prob, C.byref(rng), dev.p(d_state), dev.stream))
        hx.check(hx.lib.hx_andersen_finalize2(dev.p(d_v2), dev.p(d_f), n, A, dt, sqT, prob, C.byref(rng), dev.p(d_state2),
                                              dev.p(d_scr), sb, None, dev.stream))
        assert np.array_equal(dev.down(d_state), dev.down(d_state2))
        a, b = dev.down(d_v), dev.down(d_v2)
        assert a[:n].tobytes() == b[:n].tobytes() and a[A:A + n].tobytes() == b[A:A + n].tobytes()
    assert np.any(np.all(a[A:A + n, :3] == 0, axis=1))                # collisions happened


# ---- a14-a16 observables -----------------------------------------------------------------------------------------

def test_sample_positions_and_msd(hx, oracle, dev):
    n = 100000
    L, p, v, A = lj_fluid(n, 0.8442, 61)
    rng = np.random.default_rng(61)
    img = np.zeros((A, 4), dtype=np.float32)
    img[:n, :3] = rng.integers(-2, 3, (n, 3))
    rid = np.arange(A, dtype=np.uint32)
    rid[:n] = rng.permutation(n)
    d_p, d_i, d_rid = dev.up(p), dev.up(img), dev.up(rid)
    d_s = dev.zeros((n, 4), np.float32)
    hx.check(hx.lib.hx_sample_positions(dev.p(d_p), dev.p(d_i), dev.p(d_rid), n, hx.f3(L), dev.p(d_s), dev.stream))
    o_s = np.zeros((n, 4), dtype=np.float32)
    oracle.oracle_sample_positions(support.ptr(p), support.ptr(img), support.ptr(rid), C.c_size_t(n), support.f3(L), support.ptr(o_s))
    s0 = dev.down(d_s)
    assert s0.tobytes() == o_s.tobytes()
    # block of 6 lag samples: random walks of increasing length
    samples = [s0]
    for k in range(5):
        nxt = samples[-1].copy()
        nxt[:, :3] += rng.normal(scale=0.1 * (k + 1), size=(n, 3)).astype(np.float32)
        samples.append(nxt)
    d_samples = [dev.up(x) for x in samples]
    ptrs = np.array([t.data_ptr() for t in d_samples], dtype=np.uint64)
    d_ptrs = dev.up(ptrs)
    nl = len(samples)
    d_out = dev.zeros(3 * nl, np.float64)
    sb = hx.lib.hx_msd_scratch_bytes(nl)
    d_scr = dev.zeros(sb, np.uint8)
    hx.check(hx.lib.hx_msd_block(dev.p(d_samples[0]), dev.p(d_ptrs), nl, n, dev.p(d_out), dev.p(d_scr), sb, dev.stream))
    out = dev.down(d_out).reshape(nl, 3)
    for k in range(nl):
        o = np.zeros(3)
        oracle.oracle_msd(support.ptr(samples[0]), support.ptr(samples[k]), C.c_size_t(n), support.ptr(o))
        assert out[k, 0] == o[0] == n
        assert abs(out[k, 1] - o[1]) <= 1e-12 * max(o[1], 1e-30) + 1e-300
        assert abs(out[k, 2] - o[2]) <= 1e-9 * max(o[2], 1e-30) + 1e-300
    assert out[0, 1] == 0 and out[0, 2] == 0


def test_density_modes(hx, oracle, dev):
    n = 50000
    L, p, v, A = lj_fluid(n, 0.8442, 63)
    k0 = 2 * np.pi / L[0]
    q = np.array([[1, 0, 0], [0, 2, 0], [1, 1, 1], [3, 0, 1], [0, 0, 7], [2, 2, 0], [5, 1, 0], [1, 4, 2], [6, 6, 6], [0, 1, 0], [9, 0, 0]],
                 dtype=np.float64) * k0
    q = q.astype(np.float32)
    d_p, d_q = dev.up(p), dev.up(q)
    d_rho = dev.zeros(2 * len(q), np.float64)
    hx.check(hx.lib.hx_density_modes(dev.p(d_p), n, dev.p(d_q), len(q), dev.p(d_rho), dev.stream))
    rho = dev.down(d_rho)
    o_rho = np.zeros(2 * len(q))
    oracle.oracle_density_modes(support.ptr(p), C.c_size_t(n), support.ptr(q), len(q), support.ptr(o_rho))
    # per-term error of sincosf vs libm <= 2 ulp of O(1) terms; sums of n terms
    assert np.abs(rho - o_rho).max() < 4 * EPS_F * n ** 0.5 * 4
    # block sums are merged as fixed-point integers: the result does not depend on the order in which blocks finish
    for _ in range(5):
        hx.check(hx.lib.hx_density_modes(dev.p(d_p), n, dev.p(d_q), len(q), dev.p(d_rho), dev.stream))
        assert dev.down(d_rho).tobytes() == rho.tobytes()


def test_thermodynamics(hx, oracle, dev):
    n = 100000
    L, p, v, A = lj_fluid(n, 0.8442, 65)
    rng = np.random.default_rng(65)
    en = rng.normal(size=A).astype(np.float32)
    st = rng.normal(size=6 * A).astype(np.float32)
    d_v, d_e, d_s = dev.up(v), dev.up(en), dev.up(st)
    d_out = dev.zeros(hx.THERMO_DOUBLES, np.float64)
    hx.check(hx.lib.hx_thermodynamics(dev.p(d_v), dev.p(d_e), dev.p(d_s), A, n, dev.p(d_out), dev.stream))
    out = dev.down(d_out)
    o = np.zeros(8)
    oracle.oracle_thermodynamics(support.ptr(v), support.ptr(en), support.ptr(st), C.c_size_t(A), C.c_size_t(n), support.ptr(o))
    assert np.allclose(out[:7], o[:7], rtol=1e-12, atol=1e-9)


def test_launch_counter_and_errors(hx, dev):
    before = hx.lib.hx_launch_count()
    d = dev.zeros(1000, np.uint32)
    hx.check(hx.lib.hx_iota_u32(dev.p(d), 1000, dev.stream))
    assert hx.lib.hx_launch_count() == before + 1
    assert np.array_equal(dev.down(d), np.arange(1000, dtype=np.uint32))
    assert hx.lib.hx_device_count() >= 1
    assert hx.lib.hx_iota_u32(None, 10, dev.stream) < 0
