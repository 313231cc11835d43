"""Host-side behaviour of the Lua-module mirror (halmd_b200/api.py) that needs no GPU: potential construction
(lennard_jones.lua.in, adapters/modify.lua.in, adapters/truncate.lua.in) and its error behaviour."""
import ctypes as C

import numpy as np
import pytest

import support


@pytest.fixture(scope="module")
def mdsim():
    from halmd_b200 import mdsim
    return mdsim


KA = dict(epsilon=[[1.0, 1.5], [1.5, 0.5]], sigma=[[1.0, 0.8], [0.8, 0.88]])


@pytest.mark.parametrize("kind,trunc,h", [("sharp", 0, None), ("shifted", 1, None), ("smooth_r4", 2, 0.005), ("force_shifted", 3, None)])
def test_truncated_potential_tables_equal_oracle(mdsim, oracle, kind, trunc, h):
    """potential:truncate{type, cutoff[, h]}: the parameter tables handed to the kernels are those of the oracle's
    restatement of lennard_jones.cpp:52-70 and truncations/*.hpp."""
    kw = dict(cutoff=2.5) if h is None else dict(cutoff=2.5, h=h)
    pot = mdsim.potentials.pair.lennard_jones(**KA).truncate(kind, **kw)
    q = support.OLjParams()
    oracle.oracle_lj_init(C.byref(q), 2, 2, support.farr(np.ravel(KA["epsilon"])), support.farr(np.ravel(KA["sigma"])), trunc,
                          support.farr([2.5] * 4), C.c_float(h or 0.0))
    assert bytes(pot.params) == bytes(q)
    assert np.array_equal(pot.r_cut, (np.float32(2.5) * np.asarray(KA["sigma"], dtype=np.float32)))
    assert pot.truncation == kind and pot.species == 2


def test_default_truncation_is_shifted(mdsim):
    """truncate.lua.in:289 — the type defaults to "shifted"; Lua-style table argument."""
    pot = mdsim.potentials.pair.lennard_jones().truncate({"cutoff": 2.5})
    assert pot.truncation == "shifted" and pot.params.truncation == 1
    assert pot.params.en_cut[0] == pytest.approx(4 * (2.5 ** -12 - 2.5 ** -6), rel=1e-6)


def test_hard_core_modification(mdsim, oracle):
    """potential:modify({"hard_core", radius = ...}) (adapters/modify.lua.in): scalar radius -> scalar matrix, matrix
    radius per pair; the adapted potential is truncated like the plain one."""
    lj = mdsim.potentials.pair.lennard_jones(**KA)
    core = [[0.2, 0.1], [0.1, 0.0]]
    pot = lj.modify("hard_core", radius=core).truncate("force_shifted", cutoff=2.5)
    q = support.OLjParams()
    oracle.oracle_lj_init_hard_core(C.byref(q), 2, 2, support.farr(np.ravel(KA["epsilon"])), support.farr(np.ravel(KA["sigma"])),
                                    support.farr(np.ravel(core)), 3, support.farr([2.5] * 4), C.c_float(0))
    assert bytes(pot.params) == bytes(q) and pot.params.hard_core == 1
    scalar = lj.modify({"type": "hard_core", "radius": 0.25}).truncate("sharp", cutoff=2.5)
    assert list(scalar.params.r_core)[:4] == [np.float32(0.25) * np.float32(s) for s in np.ravel(KA["sigma"])]
    assert lj.r_core_sigma is None                        # modify returns a new potential
    plain = lj.truncate("sharp", cutoff=2.5)
    assert plain.params.hard_core == 0


def test_potential_errors(mdsim):
    lj = mdsim.potentials.pair.lennard_jones()
    with pytest.raises(ValueError):
        lj.truncate("smooth_r4", cutoff=2.5)              # smooth_r4 needs h
    with pytest.raises(ValueError):
        lj.truncate("shifted")                            # missing cutoff
    with pytest.raises(ValueError):
        lj.truncate("cubic", cutoff=2.5)                  # not built
    with pytest.raises(ValueError):
        lj.modify("soft_core", radius=0.1)                # invalid modification type
    with pytest.raises(ValueError):
        lj.modify("hard_core")                            # missing radius
    from halmd_b200.api import HxError
    with pytest.raises(HxError):
        lj.modify("hard_core", radius=3.0).truncate("shifted", cutoff=2.5)     # core radius beyond the cutoff


# ---- a16: wavevector shells (host logic of the C ABI, no device work) ---------------------------------------------------
def _pick(hx, L, q, tol, max_count, filt):
    nw = len(q)
    cap = nw * max_count
    vec = np.zeros((cap, 3))
    offs = (C.c_uint * (nw + 1))()
    f = None if filt is None else (C.c_uint * 3)(*filt)
    hx.check(hx.lib.hx_wavevector_pick((C.c_double * nw)(*q), nw, hx.d3(L), float(tol), int(max_count), f,
                                       vec.ctypes.data_as(hx.D3), cap, offs), "hx_wavevector_pick")
    return vec[:offs[nw]], list(offs)


def test_wavevector_shells_equal_the_reference_golden(hx):
    """hx_wavevector_pick against tests/golden/wavevectors.json — shells picked by the reference's own
    pick_lattice_points_from_shell (halmd/algorithm/host/pick_lattice_points.hpp:62-160, compiled in place and called as
    observables/utility/wavevector.cpp:40-123 does; scripts/make_golden_wavevectors.py): same vectors, bit for bit, in the
    same order, same shell boundaries — cubic and cuboid boxes, semilog grid, filters, a tolerance that empties shells."""
    g = support.golden("wavevectors.json")
    assert len(g["cases"]) >= 6
    for case in g["cases"]:
        vec, offs = _pick(hx, case["L"], case["wavenumbers"], case["tol"], case["max_count"], case["filter"])
        want = np.array([[float.fromhex(x) for x in v] for v in case["vectors"]]).reshape(-1, 3)
        assert offs == case["shell_offsets"], case["name"]
        assert vec.tobytes() == want.tobytes(), case["name"]


def test_wavevector_shells_equal_the_compiled_reference(hx, ref):
    """The same comparison against oracle/_ref/libhalmd_ref.so directly, on randomised boxes and grids."""
    ref.ref_wavevector_pick.restype = C.c_int
    rng = np.random.default_rng(17)
    for _ in range(12):
        L = list(rng.uniform(8.0, 40.0, 3)) if rng.random() < 0.5 else [float(rng.uniform(8.0, 40.0))] * 3
        q = sorted(rng.uniform(0.2, 9.0, int(rng.integers(1, 9))))
        tol, max_count = float(rng.choice([0.005, 0.01, 0.05])), int(rng.integers(1, 10))
        filt = None if rng.random() < 0.6 else [int(x) for x in rng.integers(0, 2, 3)]
        nw, cap = len(q), len(q) * max_count
        rvec = np.zeros((cap, 3))
        roffs = (C.c_uint * (nw + 1))()
        f = None if filt is None else (C.c_uint * 3)(*filt)
        assert ref.ref_wavevector_pick((C.c_double * nw)(*q), nw, hx.d3(L), C.c_double(tol), max_count, f,
                                       rvec.ctypes.data_as(hx.D3), cap, roffs) == 0
        if roffs[nw] == 0:
            assert hx.lib.hx_wavevector_pick((C.c_double * nw)(*q), nw, hx.d3(L), tol, max_count, f,
                                             rvec.ctypes.data_as(hx.D3), cap, roffs) < 0      # wavevector.cpp:84-87: logic_error
            continue
        vec, offs = _pick(hx, L, q, tol, max_count, filt)
        assert offs == list(roffs)
        assert vec.tobytes() == rvec[:roffs[nw]].tobytes()
