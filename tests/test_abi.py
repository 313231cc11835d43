"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls: no GPU here)."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

import support


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(support.ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(hxm?_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_all_declared_symbols(hx):
    names = declared_symbols()
    assert len(names) >= 38
    missing = [n for n in sorted(names) if not hasattr(hx.lib, n)]
    assert not missing, missing


def test_python_binding_covers_header(hx):
    names = {n for n in declared_symbols() if n.startswith("hx_")}
    assert names == set(hx.KERNEL_SYMBOLS)


def test_version_and_error_string(hx):
    assert hx.lib.hx_version() >= 100
    assert isinstance(hx.lib.hx_last_error_string(), bytes)


def test_host_side_plans_match_oracle(hx, oracle):
    """Sizing formulas are host code inside the product library; compare with the oracle's restatement."""
    for row in support.golden("lj_potential.json")["sizing_table"]["rows"]:
        L = (row["n"] / row["rho"]) ** (1 / 3)
        ncell, cl, cs = (C.c_uint * 3)(), (C.c_float * 3)(), C.c_uint()
        hx.check(hx.lib.hx_binning_plan(hx.d3([L] * 3), row["n"], row["rc"], row["skin"], row["occ_bin"], ncell, cl,
                                        C.byref(cs)))
        assert list(ncell) == [row["ncell"]] * 3 and cs.value == row["cell_size"]
        size = C.c_uint()
        hx.check(hx.lib.hx_nlist_plan(row["rc"], row["skin"], row["n"], L ** 3, row["occ_list"], cs.value, C.byref(size)))
        assert size.value == row["list_size"]
        assert hx.lib.hx_array_size(row["n"]) == row["array_size"]
        assert hx.lib.hx_hilbert_depth(hx.d3([L] * 3)) == row["depth"]
        ncell_o, cl_o, cs_o = (C.c_uint * 3)(), (C.c_float * 3)(), C.c_uint()
        oracle.oracle_binning_plan(support.d3([L] * 3), C.c_size_t(row["n"]), C.c_float(row["rc"]), C.c_float(row["skin"]),
                                   C.c_double(row["occ_bin"]), ncell_o, cl_o, C.byref(cs_o))
        assert list(cl) == list(cl_o)


def test_potential_tables_match_oracle(hx, oracle):
    eps, sigma, cut = [1.0, 1.5, 1.5, 0.5], [1.0, 0.8, 0.8, 0.88], [2.5] * 4
    for trunc, h in ((0, 0.0), (1, 0.0), (2, 0.005), (3, 0.0)):
        p = hx.LjParams()
        hx.check(hx.lib.hx_potential_lj_init(C.byref(p), 2, 2, hx.farr(eps), hx.farr(sigma), trunc, hx.farr(cut), h))
        q = support.OLjParams()
        oracle.oracle_lj_init(C.byref(q), 2, 2, support.farr(eps), support.farr(sigma), trunc, support.farr(cut), C.c_float(h))
        assert bytes(p) == bytes(q)


def test_hard_core_tables_match_oracle(hx, oracle):
    """adapters/hard_core.hpp:56-72, :86-94 under the truncations: cutoff energy and force of the adapted potential."""
    eps, sigma, cut, core = [1.0, 1.5, 1.5, 0.5], [1.0, 0.8, 0.8, 0.88], [2.5] * 4, [0.375, 0.5, 0.5, 0.0]
    for trunc, h in ((0, 0.0), (1, 0.0), (2, 0.005), (3, 0.0)):
        p = hx.LjParams()
        hx.check(hx.lib.hx_potential_lj_init_hard_core(C.byref(p), 2, 2, hx.farr(eps), hx.farr(sigma), hx.farr(core), trunc,
                                                       hx.farr(cut), h))
        q = support.OLjParams()
        oracle.oracle_lj_init_hard_core(C.byref(q), 2, 2, support.farr(eps), support.farr(sigma), support.farr(core), trunc,
                                        support.farr(cut), C.c_float(h))
        assert bytes(p) == bytes(q)
        assert p.hard_core == 1 and list(p.r_core)[:4] == [np.float32(c) * np.float32(s) for c, s in zip(core, sigma)]
    p = hx.LjParams()
    assert hx.lib.hx_potential_lj_init_hard_core(C.byref(p), 1, 1, hx.farr([1.0]), hx.farr([1.0]), hx.farr([2.5]), 1,
                                                 hx.farr([2.5]), 0.0) < 0       # core radius beyond the cutoff


def test_hoover_chain_matches_oracle(hx, oracle):
    """Host side of the Nose-Hoover thermostat in the product library against the oracle's restatement of
    verlet_nvt_hoover.cpp:79-95, :175-222: identical chain trajectories (bit for bit) for a fluctuating kinetic energy."""
    from test_oracle_properties import OHooverChain
    oracle.oracle_hoover_chain_propagate.restype = C.c_double
    oracle.oracle_hoover_chain_energy.restype = C.c_double
    n, T, fres, dt = 4000, 0.7, 3.0, 0.005
    a, b = hx.HooverChain(), OHooverChain()
    hx.check(hx.lib.hx_hoover_chain_init(C.byref(a), n, 3, T, fres))
    oracle.oracle_hoover_chain_init(C.byref(b), C.c_size_t(n), 3, C.c_double(T), C.c_double(fres))
    assert bytes(a) == bytes(b)
    rng = np.random.default_rng(8)
    ek = 3 * n * 2.0
    for _ in range(2000):
        sa = hx.lib.hx_hoover_chain_propagate(C.byref(a), ek, dt)
        sb = oracle.oracle_hoover_chain_propagate(C.byref(b), C.c_double(ek), C.c_double(dt))
        assert sa == sb
        ek = ek * sa * sa * (1 + 0.01 * rng.normal())
    assert bytes(a) == bytes(b)
    assert hx.lib.hx_hoover_chain_energy(C.byref(a), n) == oracle.oracle_hoover_chain_energy(C.byref(b), C.c_size_t(n))
    assert abs(ek / (3 * n) - T) < 0.2 * T                      # the thermostat pulled the temperature to the target
    hx.check(hx.lib.hx_hoover_chain_set_temperature(C.byref(a), 2 * T, fres))
    assert a.mass_xi[1] == pytest.approx(2 * T / (2 * np.pi * fres) ** 2, rel=1e-15) and a.en_kin_target_2 == 3 * n * 2 * T
    assert hx.lib.hx_hoover_chain_init(C.byref(a), 0, 3, T, fres) < 0
    assert hx.lib.hx_hoover_chain_set_temperature(C.byref(a), -1.0, fres) < 0


def test_planar_wall_tables_match_oracle(hx):
    """hx_planar_wall_init (host code in the product library) against the test-side restatement of
    planar_wall.cpp:83-112, with the parameters of the reference's unit test."""
    a = support.REF_WALL
    w = hx.PlanarWall()
    f32 = lambda v: hx.farr(np.asarray(v, dtype=np.float32).reshape(-1))      # noqa: E731
    hx.check(hx.lib.hx_planar_wall_init(C.byref(w), 3, 2, f32(a["offset"]), f32(a["surface_normal"]), f32(a["epsilon"]),
                                        f32(a["sigma"]), f32(a["wetting"]), f32(a["cutoff"]), a["smoothing"]))
    assert bytes(w) == bytes(support.make_oracle_wall(**a))
    assert hx.lib.hx_planar_wall_init(C.byref(w), 9, 1, f32([0] * 9), f32([[1, 0, 0]] * 9), f32([1] * 9), f32([1] * 9),
                                      f32([1] * 9), f32([1] * 9), 0.01) < 0
    assert hx.lib.hx_planar_wall_init(C.byref(w), 1, 1, f32([0]), f32([[0, 0, 0]]), f32([1]), f32([1]), f32([1]), f32([1]), 0.01) < 0


def test_contract_violations_return_error_codes(hx):
    p = hx.LjParams()
    assert hx.lib.hx_potential_lj_init(C.byref(p), 5, 5, hx.farr([1] * 25), hx.farr([1] * 25), 2, hx.farr([2.5] * 25), 0.005) < 0
    assert b"ntype" in hx.lib.hx_last_error_string()
    assert hx.lib.hx_potential_lj_init(C.byref(p), 1, 1, hx.farr([1]), hx.farr([1]), 7, hx.farr([2.5]), 0.005) < 0
    assert hx.lib.hx_shell_rows_plan(-1.0, None) < 0
    assert hx.lib.hx_displacement_exceeds(np.float32(0.0626).view(np.uint32).item(), 0.5) == 1   # sqrt(.0626) > .25
    assert hx.lib.hx_displacement_exceeds(np.float32(0.0624).view(np.uint32).item(), 0.5) == 0
