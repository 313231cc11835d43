"""Pin the oracle's pair-potential functions against the golden tables held by the reference's own unit tests
(tests/golden/lj_potential.json; sources cited there)."""
import ctypes as C

import numpy as np
import pytest

import support

EPS_D = np.finfo(np.float64).eps
EPS_F = np.finfo(np.float32).eps
TRUNC = {"sharp": 0, "shifted": 1, "smooth_r4": 2, "force_shifted": 3}


def close_fraction(a, b, tol):
    # BOOST_CHECK_CLOSE_FRACTION: |a-b| <= tol * min(|a|,|b|) (strong check); exact zeros must match
    if a == 0 or b == 0:
        return abs(a - b) <= tol
    return abs(a - b) <= tol * min(abs(a), abs(b))


def pair_d(oracle, eps, sigma, cutoff, trunc, h, r):
    f, e = C.c_double(), C.c_double()
    oracle.oracle_pair_d(eps, sigma, cutoff, trunc, h, r * r, C.byref(f), C.byref(e))
    return f.value, e.value


@pytest.mark.parametrize("trunc", ["sharp", "shifted", "smooth_r4", "force_shifted"])
def test_lj_tables_host_double(oracle, trunc):
    """lennard_jones.cpp:275-356 — host functor, tolerance 5 eps_double."""
    g = support.golden("lj_potential.json")
    lj = g["lennard_jones"]
    for pair, par in g["pairs"].items():
        for r, fval, en in lj[trunc][pair]:
            cutoff_r = lj["cutoff_sigma"] * par["sigma"]
            f, e = pair_d(oracle, par["epsilon"], par["sigma"], lj["cutoff_sigma"], TRUNC[trunc], lj["h"], r)
            if r >= cutoff_r:
                # the functor is only evaluated inside the cutoff by its callers; tables list 0 there
                if trunc != "sharp" and fval == 0 and en == 0:
                    continue
            assert close_fraction(f, fval, 5 * EPS_D), (trunc, pair, r, f, fval)
            assert close_fraction(e, en, 5 * EPS_D) or abs(e - en) < 1e-19, (trunc, pair, r, e, en)


def test_smooth_r4_near_cutoff_double(oracle):
    """smooth_r4.cpp:48-142 — tolerance k eps (1 + rc / (rc - r))."""
    g = support.golden("lj_potential.json")
    t = g["smooth_r4_near_cutoff"]
    wca = float(np.float32(2.0 ** (1 / 6.0)))
    cut = {"aa": 2.5, "ab": 2.0, "bb": wca}
    k = {"aa": 8, "ab": 8, "bb": 9}
    for pair, par in g["pairs"].items():
        rc = cut[pair] * par["sigma"]
        for r, fval, en in t[pair]:
            f, e = pair_d(oracle, par["epsilon"], par["sigma"], cut[pair], 2, t["h"], r)
            tol = k[pair] * EPS_D * (1 + rc / (rc - r))
            assert close_fraction(f, fval, tol), (pair, r, f, fval)
            assert close_fraction(e, en, tol), (pair, r, e, en)


def make_params(oracle, eps, sigma, cutoff, trunc, h):
    p = support.OLjParams()
    n = int(round(len(eps) ** 0.5))
    oracle.oracle_lj_init(C.byref(p), n, n, support.farr(eps), support.farr(sigma), trunc, support.farr(cutoff),
                          C.c_float(h))
    return p


@pytest.mark.parametrize("trunc", ["sharp", "shifted", "smooth_r4", "force_shifted"])
def test_lj_tables_gpu_functor_float(oracle, trunc):
    """The float (GPU-semantics) functor against the same tables: lennard_jones.cpp:392-436 compares the GPU
    kernel with the host functor at 10 eps_float; smooth_r4 adds 8 eps (1 + rc/(rc-r)) (smooth_r4.cpp:229)."""
    g = support.golden("lj_potential.json")
    lj = g["lennard_jones"]
    eps = [1.0, 0.5, 0.5, 0.25]
    sigma = [1.0, 2.0, 2.0, 4.0]
    p = make_params(oracle, eps, sigma, [lj["cutoff_sigma"]] * 4, TRUNC[trunc], lj["h"])
    idx = {"aa": 0, "ab": 1, "bb": 3}
    for pair, par in g["pairs"].items():
        rc = lj["cutoff_sigma"] * par["sigma"]
        for r, fval, en in lj[trunc][pair]:
            f, e = C.c_float(), C.c_float()
            rr = np.float32(r) * np.float32(r)
            inside = oracle.oracle_pair_f(C.byref(p), idx[pair], C.c_float(rr), C.byref(f), C.byref(e))
            if not inside:
                assert r >= rc * (1 - 1e-6)
                continue
            tol = 10 * EPS_F
            if trunc == "smooth_r4":
                tol += 8 * EPS_F * (1 + rc / (rc - r))
            if trunc == "force_shifted":
                tol = 5e2 * 10 * EPS_F         # lennard_jones.cpp:266-271: the subtraction is ill-conditioned
            # r^2 in float carries 1 ulp; r^-14 amplifies it 7-fold
            tol += 8 * EPS_F
            assert abs(f.value - fval) <= tol * max(abs(fval), 1e-30), (trunc, pair, r, f.value, fval)
            # near the zero of the shifted potential only an absolute bound is meaningful
            assert abs(e.value - en) <= tol * max(abs(en), abs(en - e.value) * 0 + 4 * par["epsilon"] * (par["sigma"] / r) ** 12,
                                                   1e-30), (trunc, pair, r, e.value, en)


def test_hard_core_adapter_closed_form(oracle):
    """adapters/hard_core_kernel.hpp:66-76 over lennard_jones, sharp truncation: U(r) = 4 eps [(sigma / (r - r_core))^12
    - (sigma / (r - r_core))^6], fval = -U'(r) / r, evaluated in double; tolerance of the reference's adapter test
    (power_law_hard_core.cpp:155-160: eps * n * (1 + r / (r - r_core)), n = 12 here, float functor)."""
    eps, sigma, core, cut = [1.0, 0.5, 0.5, 0.25], [1.0, 2.0, 2.0, 4.0], [0.375, 0.5, 0.5, 0.75], [5.0, 6.0, 6.0, 7.0]
    p = support.OLjParams()
    oracle.oracle_lj_init_hard_core(C.byref(p), 2, 2, support.farr(eps), support.farr(sigma), support.farr(core), 0,
                                    support.farr(cut), C.c_float(0))
    assert p.hard_core == 1
    for idx in (0, 1, 3):
        e_, s_, rc_ = eps[idx], sigma[idx], core[idx] * sigma[idx]
        assert p.r_core[idx] == np.float32(core[idx]) * np.float32(sigma[idx])
        for x in (0.5, 0.8, 1.0, 1.12, 1.5, 2.0, 3.0, 4.5):
            r = rc_ + x * s_
            f, e = C.c_float(), C.c_float()
            rr = np.float32(r) * np.float32(r)
            assert oracle.oracle_pair_f(C.byref(p), idx, C.c_float(rr), C.byref(f), C.byref(e))
            r_ = float(np.sqrt(np.float64(rr)))
            q6 = (s_ / (r_ - rc_)) ** 6
            en = 4 * e_ * (q6 * q6 - q6)
            fval = 24 * e_ * (2 * q6 * q6 - q6) / (r_ - rc_) / r_
            tol = 14 * EPS_F * (1 + r / (r - rc_))
            # (absolute scale of the two competing terms: the force changes sign at 2^(1/6) sigma)
            assert abs(f.value - fval) <= tol * max(abs(fval), 24 * e_ * q6 / (r_ - rc_) / r_), (idx, x, f.value, fval)
            assert abs(e.value - en) <= tol * max(abs(en), 4 * e_ * q6), (idx, x, e.value, en)
    # shifted / force-shifted tables use the adapted potential at the cutoff (shifted.hpp, force_shifted.hpp:69-70)
    oracle.oracle_lj_init_hard_core(C.byref(p), 2, 2, support.farr(eps), support.farr(sigma), support.farr(core), 3,
                                    support.farr(cut), C.c_float(0))
    rc_, r_c = core[0] * sigma[0], cut[0] * sigma[0]
    q6 = (sigma[0] / (r_c - rc_)) ** 6
    assert abs(p.en_cut[0] - 4 * eps[0] * (q6 * q6 - q6)) < 1e-5 * abs(4 * eps[0] * q6)
    assert abs(p.force_cut[0] - 24 * eps[0] * (2 * q6 * q6 - q6) / (r_c - rc_)) < 1e-5 * abs(24 * eps[0] * q6 / (r_c - rc_))
    f, e = C.c_float(), C.c_float()
    assert not oracle.oracle_pair_f(C.byref(p), 0, C.c_float(np.float32(r_c) ** 2), C.byref(f), C.byref(e))
    assert oracle.oracle_pair_f(C.byref(p), 0, C.c_float(np.float32(r_c * (1 - 1e-4)) ** 2), C.byref(f), C.byref(e))
    assert abs(f.value) < 1e-3 * abs(p.force_cut[0]) and abs(e.value) < 1e-6     # force and energy vanish at the cutoff


def test_lj_init_tables(oracle):
    """Derived tables as in lennard_jones.cpp:52-70 / smooth_r4.hpp:55-91, KA mixture parameters."""
    eps = [1.0, 1.5, 1.5, 0.5]
    sigma = [1.0, 0.8, 0.8, 0.88]
    p = make_params(oracle, eps, sigma, [2.5] * 4, 2, 0.005)
    assert p.rri_smooth == np.float32(0.005 ** -2)
    for i in range(4):
        s = np.float32(sigma[i])
        assert p.sigma2[i] == s * s
        assert p.r_cut[i] == s * np.float32(2.5)
        assert p.rr_cut[i] == np.float32(p.r_cut[i]) * np.float32(p.r_cut[i])
        u = 4 * eps[i] * ((1 / 2.5) ** 12 - (1 / 2.5) ** 6)
        assert abs(p.en_cut[i] - u) < 1e-6 * abs(u)


def test_planar_wall_oracle_against_reference_test_formula(oracle):
    """The reference's own unit test (test/unit/mdsim/potentials/external/planar_wall.cpp:149-201): three walls, a
    binary mixture, seven test positions, force and energy against the closed formula of that test.  The reference
    checks its float GPU path to 10 float epsilon; the wall-2 / species-0 entry has cutoff 10 and smoothing 0.01, where
    float rounding of (|d| - cutoff) / smoothing is amplified, hence the absolute term.  The formula is evaluated at the
    float-rounded positions the kernel sees."""
    w = support.make_oracle_wall(**support.REF_WALL)
    L = np.array([100.0, 100.0, 100.0])
    for species in range(2):
        pos = np.zeros((len(support.REF_WALL_POSITIONS), 4), dtype=np.float32)
        pos[:, :3] = support.REF_WALL_POSITIONS
        pos[:, 3] = np.array([species] * len(pos), dtype=np.uint32).view(np.float32)
        f = np.zeros_like(pos)
        e = np.zeros(len(pos), dtype=np.float32)
        st = np.ones(6 * len(pos), dtype=np.float32)
        oracle.oracle_planar_wall(C.byref(w), support.ptr(pos), C.c_size_t(len(pos)), support.f3(L), 1, support.ptr(f),
                                  support.ptr(e), support.ptr(st), C.c_size_t(len(pos)))
        assert np.all(st == 0)
        for k, r in enumerate(support.REF_WALL_POSITIONS):
            f2, e2 = np.zeros(3), 0.0
            for i in range(3):
                n = np.asarray(support.REF_WALL["surface_normal"][i], dtype=np.float64)
                fi, ei = support.reference_test_wall(pos[k, :3].astype(np.float64), n / np.linalg.norm(n),
                                                     support.REF_WALL["offset"][i], support.REF_WALL["epsilon"][i][species],
                                                     support.REF_WALL["sigma"][i][species], support.REF_WALL["wetting"][i][species],
                                                     support.REF_WALL["cutoff"][i][species], support.REF_WALL["smoothing"])
                f2 += fi
                e2 += ei
            tol = 10 * 1.1920929e-07
            assert np.all(np.abs(f[k, :3] - f2) <= tol * np.abs(f2).max() + 2e-6 * abs(e2) + 1e-12), (species, k, f[k, :3], f2)
            assert abs(e[k] - e2) <= tol * abs(e2) + 1e-9, (species, k, e[k], e2)


def test_blocking_schedule_golden():
    """tests/golden/dynamics.json (scripts/make_golden_dynamics.py: the reference's clock, blocking_scheme, sample buffers and
    correlation<> compiled in place and driven like sampler:run): the schedule restatement the GPU tests hold the product's
    blocking scheme to reproduces the recorded intervals and the sequence of correlated sample pairs."""
    g = support.golden("dynamics.json")
    assert len(g["blocking"]) >= 4
    for case in g["blocking"]:
        dt, max_lag, res, size, shift, sep, start, steps = case["args"]
        interval, events = support.blocking_reference(int(steps), dt, max_lag, res, int(size), int(shift), int(sep), start=int(start))
        assert interval == case["intervals"]
        mine = support.blocking_events(events)
        assert np.array_equal(mine.reshape(-1), np.array(case["events"], dtype=np.int64))
        cnt = np.zeros((len(interval), int(size)), dtype=np.int64)
        for lev, first, second in mine:
            cnt[lev, (second - first) // interval[lev]] += 1
        assert np.array_equal(cnt, np.array(case["count"]))


def test_isf_product_golden():
    """The product's host-side ISF product (hx_isf_accumulate, intermediate_scattering_function.hpp:90-113) and wavevector
    picker against the accumulators and shells recorded from the reference's own classes (tests/golden/dynamics.json), bit
    for bit in count and mean."""
    import sys
    import os
    sys.path.insert(0, os.path.join(support.ROOT, "scripts"))
    from make_golden_dynamics import isf_samples
    from halmd_b200 import _lib as hx
    g = support.golden("dynamics.json")["isf"]
    L, q = np.array(g["L"]), np.array(g["wavenumber"])
    cap = len(q) * g["max_count"]
    vec = np.zeros((cap, 3))
    offs = np.zeros(len(q) + 1, dtype=np.uint32)
    hx.check(hx.lib.hx_wavevector_pick(q.ctypes.data_as(hx.D3), len(q), L.ctypes.data_as(hx.D3), g["tol"], g["max_count"], None,
                                       vec.ctypes.data_as(hx.D3), cap, offs.ctypes.data_as(hx.U3)))
    shells = [[int(offs[i]), int(offs[i + 1])] for i in range(len(q)) if offs[i + 1] > offs[i]]
    assert shells == g["shells"] and shells[-1][1] == g["nq"]
    assert [float(q[i]) for i in range(len(q)) if offs[i + 1] > offs[i]] == g["kept_wavenumbers"]
    shell_offsets = np.array([s[0] for s in shells] + [shells[-1][1]], dtype=np.uint32)
    rho1, rho2 = isf_samples(g["nq"], g["nsamples"], g["seed"])
    acc = np.zeros((len(shells), 3))
    for s in range(g["nsamples"]):
        a, b = np.ascontiguousarray(rho1[s]), np.ascontiguousarray(rho2[s])
        hx.check(hx.lib.hx_isf_accumulate(a.ctypes.data_as(hx.D3), b.ctypes.data_as(hx.D3), shell_offsets.ctypes.data_as(hx.U3),
                                          len(shells), g["norm"], acc.ctypes.data_as(hx.D3)))
    ref = np.array([[float.fromhex(x) for x in row] for row in g["result_hex"]])
    assert np.array_equal(acc[:, 0], ref[:, 0])
    assert acc[:, 1].tobytes() == ref[:, 1].tobytes()
    assert np.allclose(acc[:, 2] / acc[:, 0], ref[:, 2], rtol=1e-15, atol=0)
