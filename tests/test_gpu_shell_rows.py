"""Private shell-ordered rows of the pair-force kernel (hx_shell_rows_build / hx_pair_force_shells, include/halmd_b200.h).

The public neighbour list stays the reference's (tests/test_gpu_kernels.py, tests/test_gpu_reference_kernels.py); here the
private copy is checked to be a stable regrouping of it by build-time distance, and the force kernel that walks only
the shells which can reach the cutoff is checked against the full walk of the same list (hx_pair_force), the CPU oracle
(pair_trunc_kernel.cuh:159-255 restated) and, in tests/test_gpu_reference_kernels.py, the reference's own kernel."""
import ctypes as C

import numpy as np
import pytest

import support
from test_gpu_kernels import Dev, build_lists, check_force, ORACLE_TOL  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(cuda):
    return Dev(cuda)


def make_shells(hx, dev, s, par, skin):
    """hx_count_max -> pitch -> hx_shell_rows_build on the lists of `s` (positions at build time)."""
    n = s["n"]
    d_count = dev.up(s["count"].astype(np.uint32))
    d_max = dev.zeros(1, np.uint32)
    hx.check(hx.lib.hx_count_max(dev.p(d_count), n, dev.p(d_max), dev.stream), "hx_count_max")
    longest = int(dev.down(d_max)[0])
    assert longest == int(s["count"].max())
    pitch = (longest + 32) // 32 * 32
    sh = hx.ShellRows()
    hx.check(hx.lib.hx_shell_rows_plan(float(skin), C.byref(sh)))
    words = int(hx.lib.hx_shell_rows_words(n, pitch))
    assert words >= (n + 7) // 8 * 8 * pitch + 256
    d_rows = dev.up(np.full(words, 0xDEADBEEF, dtype=np.uint32))
    d_ends = dev.zeros(n * hx.NSHELL, np.uint16)
    sh.rows_dev, sh.ends_dev, sh.pitch = d_rows.data_ptr(), d_ends.data_ptr(), pitch
    hx.check(hx.lib.hx_shell_rows_build(C.byref(par), dev.p(s["d_nl"]), s["stride"], dev.p(d_count), dev.p(s["d_pos"]),
                                        dev.p(s["d_pos"]), n, hx.f3(s["L"]), C.byref(sh), dev.stream), "hx_shell_rows_build")
    return sh, d_rows, d_ends, pitch


def decode_rows(flat, n, pitch):
    """Private rows from the tile-interleaved layout documented in include/halmd_b200.h (hx_shell_rows)."""
    i = np.arange(n, dtype=np.int64)[:, None]
    k = np.arange(pitch, dtype=np.int64)[None, :]
    e = k % 8
    slot = np.where(e < 4, 2 * e, 2 * (e - 4) + 1)
    return flat[(i // 8) * 8 * pitch + (k // 8) * 64 + (i % 8) * 8 + slot]


def lj(hx, oracle, nt, eps, sigma, trunc, cut, h=0.005):
    par = hx.LjParams()
    hx.check(hx.lib.hx_potential_lj_init(C.byref(par), nt, nt, hx.farr(eps), hx.farr(sigma), trunc, hx.farr(cut), h))
    opar = support.OLjParams()
    oracle.oracle_lj_init(C.byref(opar), nt, nt, support.farr(eps), support.farr(sigma), trunc, support.farr(cut), C.c_float(h))
    return par, opar


@pytest.mark.parametrize("n,rho,rc,skin,ob,ol,frac,eps,sigma", [
    (16000, 0.8442, [2.5], 0.5, 0.5, 0.4, None, [1.0], [1.0]),
    (32000, 1.2, [2.5, 2.0, 2.0, 2.2], 0.3, 0.7, 0.7, 0.8, [1.0, 1.5, 1.5, 0.5], [1.0, 0.8, 0.8, 0.88]),
])
def test_shell_rows_are_a_stable_regrouping(hx, oracle, dev, n, rho, rc, skin, ob, ol, frac, eps, sigma):
    s = build_lists(hx, oracle, dev, n, rho, 71, rc, skin, ob, ol, frac, layout=1, fill_all=0, jitter=0.08 if frac else 0.12)
    nt = s["nt"]
    par, _ = lj(hx, oracle, nt, eps, sigma, 2, [2.5] * (nt * nt))
    sh, d_rows, d_ends, pitch = make_shells(hx, dev, s, par, skin)
    flat = dev.down(d_rows)
    assert np.all(flat[(n + 7) // 8 * 8 * pitch:] == 0)       # slack behind the last tile
    rows = decode_rows(flat, n, pitch)
    ends = dev.down(d_ends).reshape(n, hx.NSHELL).astype(np.int64)
    cnt = s["count"].astype(np.int64)
    pub = s["nl"].reshape(n, s["stride"])
    assert np.all(np.diff(ends, axis=1) >= 0)
    assert np.array_equal(ends[:, -1], cnt)
    k = np.arange(pitch)[None, :]
    # behind its entries a row holds the particle's own index as far as the force kernel walks the rows of its tile of 8
    # particles (the longest of them, in steps of 16 entries); every word up to there is initialised
    padded = np.concatenate([cnt, np.zeros((-n) % 8, dtype=np.int64)])
    longest = np.repeat(padded.reshape(-1, 8).max(axis=1), 8)[:n]
    fill_end = np.minimum((longest + 15) // 16 * 16, pitch)[:, None]
    fill = (k >= cnt[:, None]) & (k < fill_end)
    assert np.all(rows[fill] == np.broadcast_to(np.arange(n)[:, None], (n, pitch))[fill])
    assert not np.any(rows[k < fill_end] == 0xDEADBEEF)
    delta = np.array(list(sh.delta), dtype=np.float64)
    assert abs(delta[-1] - skin) < 1e-6
    pos = s["p"][:n, :3].astype(np.float64)
    species = support.species_of(s["p"], s["A"])[:n].astype(np.int64)
    r_cut = np.array([par.r_cut[i] for i in range(nt * nt)], dtype=np.float64)
    L = np.asarray(s["L"], dtype=np.float64)
    rng = np.random.default_rng(5)
    for i in rng.choice(n, size=400, replace=False):
        c = cnt[i]
        mine, theirs = rows[i, :c], pub[i, :c]
        assert np.array_equal(np.sort(mine), np.sort(theirs))
        d = pos[i] - pos[mine]
        d -= np.rint(d / L) * L
        r = np.sqrt((d ** 2).sum(axis=1))
        x = r - r_cut[species[i] * nt + species[mine]]         # distance beyond the pair's cutoff at build time
        order_in_public = {int(j): q for q, j in enumerate(theirs)}
        lo = 0
        for sidx in range(hx.NSHELL):
            hi = ends[i, sidx]
            seg = x[lo:hi]
            if sidx > 0:
                assert np.all(seg > delta[sidx - 1] - 1e-5)
            if sidx < hx.NSHELL - 1:
                assert np.all(seg <= delta[sidx] + 1e-5)
            q = [order_in_public[int(j)] for j in mine[lo:hi]]
            assert q == sorted(q)                              # stable: the reference's order within a shell
            lo = hi


def displaced(s, rng, dmax):
    """Positions moved by up to dmax (most particles far less), reduced to the primary cell; float32 high words."""
    n, A = s["n"], s["A"]
    p0 = s["p"][:n, :3].astype(np.float64)
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    amp = dmax * rng.random(n) ** 3
    amp[rng.integers(n)] = dmax
    p1 = p0 + u * amp[:, None]
    L = np.asarray(s["L"], dtype=np.float64)
    p1 -= np.rint(p1 / L) * L
    p = s["p"].copy()
    p[:n, :3] = p1.astype(np.float32)
    # maximum squared displacement as the integrator's reduction computes it (float, minimum image)
    d = p[:n, :3].astype(np.float32) - s["p"][:n, :3].astype(np.float32)
    Lf = L.astype(np.float32)
    d = d - np.rint(d / Lf) * Lf
    rr = (d.astype(np.float64) ** 2).sum(axis=1).max()
    return p, np.float32(rr * (1 + 1e-6))


def shell_force(hx, dev, s, par, sh, d_pos, d_r0, ctl, skin, lanes=4, flags=0, force_zero=1, f0=None):
    n, A = s["n"], s["A"]
    d_f = dev.up(f0) if f0 is not None else dev.zeros((A, 4), np.float32)
    d_ctl = dev.up(ctl) if ctl is not None else None
    hx.check(hx.lib.hx_pair_force_shells(C.byref(par), dev.p(d_pos), dev.p(d_pos), C.byref(sh), dev.p(d_r0) if d_r0 is not None else None,
                                         n, hx.f3(s["L"]), flags | hx.FORCE_LANES(lanes), force_zero, dev.p(d_f),
                                         dev.p(d_ctl) if d_ctl is not None else None, float(np.float32(skin) / 2), dev.stream),
             "hx_pair_force_shells")
    return dev.down(d_f), (dev.down(d_ctl) if d_ctl is not None else None)


def full_force(hx, dev, s, par, d_pos, lanes=4, flags=0):
    n, A = s["n"], s["A"]
    d_f = dev.zeros((A, 4), np.float32)
    hx.check(hx.lib.hx_pair_force(C.byref(par), dev.p(d_pos), dev.p(d_pos), dev.p(s["d_nl"]), s["size"], s["stride"],
                                  s["layout"] | flags | hx.FORCE_LANES(lanes), n, hx.f3(s["L"]), 1, 0.5, dev.p(d_f), None, None, A,
                                  dev.stream), "hx_pair_force")
    return dev.down(d_f)


def oracle_force(oracle, s, opar, p):
    n, A = s["n"], s["A"]
    of = np.zeros((A, 4), dtype=np.float32)
    oracle.oracle_pair_force(C.byref(opar), support.ptr(p), support.ptr(p), support.ptr(s["o_nl"]), s["size"],
                             C.c_size_t(s["stride"]), s["layout"], C.c_size_t(n), support.f3(s["L"]), 1, C.c_float(0.5),
                             support.ptr(of), None, None, C.c_size_t(A))
    return of


def close(f, g, n, tol):
    d = np.linalg.norm(f[:n, :3].astype(np.float64) - g[:n, :3].astype(np.float64), axis=1)
    mag = np.linalg.norm(g[:n, :3].astype(np.float64), axis=1)
    rel = d / np.maximum(mag, np.sqrt((mag ** 2).mean()))
    assert rel.max() < tol, (rel.max(), int(rel.argmax()))
    return rel


@pytest.mark.parametrize("n,trunc,dmax", [(16000, t, d) for d in (0.0, 0.03, 0.12, 0.249) for t in (0, 1, 2, 3)] + [(15997, 2, 0.12)])
def test_shell_force_equals_full_walk(hx, oracle, dev, n, trunc, dmax):
    """Lists built at r0, positions moved by up to dmax: the prefix walk must not lose a pair inside the cutoff
    (same forces as the full walk up to the order of the compensated sums) and match the oracle.  n = 15997: the last
    tile of 8 particles is incomplete (exercised for smooth_r4 at one displacement)."""
    skin = 0.5
    lanes = 4
    s = build_lists(hx, oracle, dev, n, 0.8442, 73, [2.5], skin, 0.5, 0.4, layout=1, fill_all=0, reduce=True)
    par, opar = lj(hx, oracle, 1, [1.0], [1.0], trunc, [2.5])
    sh, d_rows, d_ends, pitch = make_shells(hx, dev, s, par, skin)
    p, max_rr = displaced(s, np.random.default_rng(7), dmax)
    d_pos = dev.up(p)
    n = s["n"]
    ctl = np.array([max_rr.view(np.uint32), 0, 0, 0], dtype=np.uint32)
    f, ctl_out = shell_force(hx, dev, s, par, sh, d_pos, s["d_pos"], ctl, skin, lanes=lanes)
    assert ctl_out[1] == 0 and ctl_out[2] == 1                 # not halted, one step counted
    g = full_force(hx, dev, s, par, d_pos, lanes=lanes)
    rel = close(f, g, n, 5e-7)
    assert np.median(rel) == 0.0                               # same pairs, same arithmetic: typically identical bits
    close(f, oracle_force(oracle, s, opar, p), n, ORACLE_TOL)
    # the reduced-positions fast path changes nothing; without r0 / ctl the whole row is walked
    f2, _ = shell_force(hx, dev, s, par, sh, d_pos, s["d_pos"], ctl, skin, lanes=lanes, flags=hx.FORCE_POSITIONS_REDUCED)
    assert f2.tobytes() == f.tobytes()
    f3, _ = shell_force(hx, dev, s, par, sh, d_pos, None, None, skin, lanes=lanes)
    close(f3, g, n, 5e-7)


def test_shell_force_walks_a_prefix_only(hx, oracle, dev):
    """With an understated displacement bound the kernel does drop pairs: evidence that the prefix is what is walked
    (and that the bound of the previous test is the one that matters)."""
    skin = 0.5
    s = build_lists(hx, oracle, dev, 16000, 0.8442, 73, [2.5], skin, 0.5, 0.4, layout=1, fill_all=0, reduce=True)
    par, opar = lj(hx, oracle, 1, [1.0], [1.0], 0, [2.5])      # sharp truncation: a lost pair at the cutoff is visible
    sh, *_keep = make_shells(hx, dev, s, par, skin)
    p, max_rr = displaced(s, np.random.default_rng(7), 0.249)
    d_pos = dev.up(p)
    g = full_force(hx, dev, s, par, d_pos)
    lie = np.array([np.float32(1e-8).view(np.uint32), 0, 0, 0], dtype=np.uint32)
    f, _ = shell_force(hx, dev, s, par, sh, d_pos, d_pos, lie, skin)      # r0 = current positions: d_i = 0
    d = np.linalg.norm(f[:s["n"], :3].astype(np.float64) - g[:s["n"], :3].astype(np.float64), axis=1)
    assert (d > 1e-4).sum() > 10


def test_shell_force_binary_mixture(hx, oracle, dev):
    """Kob-Andersen table: shells are relative to the cutoff of each species pair."""
    eps, sigma = [1.0, 1.5, 1.5, 0.5], [1.0, 0.8, 0.8, 0.88]
    rc = [2.5 * x for x in sigma]
    skin = 0.3
    s = build_lists(hx, oracle, dev, 32000, 1.2, 75, rc, skin, 0.7, 0.7, species_frac=0.8, layout=1, fill_all=0, jitter=0.08)
    par, opar = lj(hx, oracle, 2, eps, sigma, 2, [2.5] * 4)
    sh, *_keep = make_shells(hx, dev, s, par, skin)
    for dmax in (0.02, 0.149):
        p, max_rr = displaced(s, np.random.default_rng(9), dmax)
        d_pos = dev.up(p)
        ctl = np.array([max_rr.view(np.uint32), 0, 0, 0], dtype=np.uint32)
        f, _ = shell_force(hx, dev, s, par, sh, d_pos, s["d_pos"], ctl, skin)
        close(f, full_force(hx, dev, s, par, d_pos), s["n"], 5e-7)
        close(f, oracle_force(oracle, s, opar, p), s["n"], ORACLE_TOL)


def test_shell_force_guard_and_accumulate(hx, oracle, dev):
    skin = 0.5
    s = build_lists(hx, oracle, dev, 4000, 0.8442, 77, [2.5], skin, 0.5, 0.4, layout=1, fill_all=0)
    par, opar = lj(hx, oracle, 1, [1.0], [1.0], 2, [2.5])
    sh, *_keep = make_shells(hx, dev, s, par, skin)
    n, A = s["n"], s["A"]
    rng = np.random.default_rng(3)
    f0 = np.zeros((A, 4), dtype=np.float32)
    f0[:n, :3] = rng.normal(size=(n, 3))
    # displacement beyond skin / 2: the kernel raises the halted flag and leaves the force array alone
    ctl = np.array([np.float32(0.0626).view(np.uint32), 0, 5, 0], dtype=np.uint32)
    f, c = shell_force(hx, dev, s, par, sh, s["d_pos"], s["d_pos"], ctl, skin, f0=f0)
    assert c[1] == 1 and c[2] == 5 and f.tobytes() == f0.tobytes()
    # already halted: untouched as well
    ctl = np.array([0, 1, 5, 0], dtype=np.uint32)
    f, c = shell_force(hx, dev, s, par, sh, s["d_pos"], s["d_pos"], ctl, skin, f0=f0)
    assert c[1] == 1 and c[2] == 5 and f.tobytes() == f0.tobytes()
    # force_zero = 0 adds to the existing force (pair_trunc_kernel.cuh:241-248)
    ctl = np.array([0, 0, 0, 0], dtype=np.uint32)
    f, c = shell_force(hx, dev, s, par, sh, s["d_pos"], s["d_pos"], ctl, skin, force_zero=0, f0=f0)
    g = full_force(hx, dev, s, par, s["d_pos"])
    g[:n, :3] += f0[:n, :3]
    close(f, g, n, 1e-6)


def test_shell_rows_256k(hx, oracle, dev):
    """BASELINE config 2 size: fraction of the stored entries walked at a mid-interval displacement."""
    skin = 0.5
    s = build_lists(hx, oracle, dev, 256000, 0.8442, 79, [2.5], skin, 0.5, 0.4, layout=1, fill_all=0, reduce=True)
    par, opar = lj(hx, oracle, 1, [1.0], [1.0], 2, [2.5])
    sh, d_rows, d_ends, pitch = make_shells(hx, dev, s, par, skin)
    p, max_rr = displaced(s, np.random.default_rng(11), 0.12)
    d_pos = dev.up(p)
    ctl = np.array([max_rr.view(np.uint32), 0, 0, 0], dtype=np.uint32)
    f, _ = shell_force(hx, dev, s, par, sh, d_pos, s["d_pos"], ctl, skin, flags=hx.FORCE_POSITIONS_REDUCED)
    close(f, full_force(hx, dev, s, par, d_pos, flags=hx.FORCE_POSITIONS_REDUCED), s["n"], 5e-7)
    close(f, oracle_force(oracle, s, opar, p), s["n"], ORACLE_TOL)
