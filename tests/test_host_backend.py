"""CPU baseline (oracle/host_backend.c = restatement of halmd/mdsim/host/**): physics checks."""
import ctypes as C

import numpy as np

import support


def make_sim(oracle, n, rho, trunc, nthreads=1, dt=0.001, seed=1, T=1.2):
    L = (n / rho) ** (1 / 3)
    rng = np.random.default_rng(seed)
    pos = support.fcc_positions(n, [L] * 3)
    vel = rng.normal(scale=np.sqrt(T), size=(n, 3))
    vel -= vel.mean(axis=0)
    sim = oracle.host_sim_create(C.c_size_t(n), support.d3([L] * 3), 1, support.darr([1.0]), support.darr([1.0]),
                                 support.darr([2.5]), trunc, C.c_double(0.005), C.c_double(0.5), C.c_double(dt), nthreads)
    assert sim
    pos = np.ascontiguousarray(pos)
    vel = np.ascontiguousarray(vel)
    oracle.host_sim_set(C.c_void_p(sim), support.ptr(pos), support.ptr(vel), None, None)
    return sim, L


def thermo(oracle, sim):
    out = np.zeros(6)
    oracle.host_sim_thermodynamics(C.c_void_p(sim), support.ptr(out))
    return out


def test_energy_conservation_smooth_r4(oracle):
    """thermodynamics.cpp:150-161 gate (scaled): |dE|/|E| small over an NVE run with the smooth cutoff."""
    sim, L = make_sim(oracle, 1372, 0.8442, 2)
    oracle.host_sim_run(C.c_void_p(sim), C.c_size_t(300))      # melt a little
    e0 = thermo(oracle, sim)
    oracle.host_sim_run(C.c_void_p(sim), C.c_size_t(500))
    e1 = thermo(oracle, sim)
    E0, E1 = e0[0] + e0[1], e1[0] + e1[1]
    assert abs(E1 - E0) / abs(E0) < 2e-5
    assert np.abs(e1[3:]).max() < 1e-12                        # centre-of-mass velocity stays zero
    assert oracle.host_sim_rebuilds(C.c_void_p(sim)) >= 2
    oracle.host_sim_destroy(C.c_void_p(sim))


def test_openmp_matches_serial(oracle):
    sim1, _ = make_sim(oracle, 1372, 0.8442, 1)
    sim4, _ = make_sim(oracle, 1372, 0.8442, 1, nthreads=4)
    for s in (sim1, sim4):
        oracle.host_sim_run(C.c_void_p(s), C.c_size_t(50))
    p1, p4 = np.zeros((1372, 3)), np.zeros((1372, 3))
    oracle.host_sim_get(C.c_void_p(sim1), support.ptr(p1), None, None, None)
    oracle.host_sim_get(C.c_void_p(sim4), support.ptr(p4), None, None, None)
    assert np.abs(p1 - p4).max() < 1e-9
    oracle.host_sim_destroy(C.c_void_p(sim1))
    oracle.host_sim_destroy(C.c_void_p(sim4))
