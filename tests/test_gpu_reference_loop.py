"""The product's step loop against the REFERENCE's own GPU step loop (oracle/ref_gpu/refsim.cpp: the reference's kernels —
verlet, max_displacement, hilbert map, radix sort, rearrange, binning, from_binning, pair_trunc — compiled in place
and launched in the order and with the synchronisations of its host classes), from identical initial conditions:
short-time trajectory, rebuild schedule, and the energy drift over 10^5 NVE steps (north star: "no worse than the
reference"; gate formula of test/integration/validation/thermodynamics.cpp:150-161).  Also the thermodynamic sums against
the reference's reduction kernels (observables/gpu/thermodynamics_kernel.cu)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import support

sys.path.insert(0, os.path.join(support.ROOT, "oracle"))
import refsim  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(cuda):
    from halmd_b200 import mdsim, simulation
    return mdsim, simulation


@pytest.fixture(scope="module")
def refloop(cuda):
    if not refsim.available():
        pytest.skip("oracle/_ref/libhalmd_ref_gpu.so not built (reference sources absent at build time)")
    return refsim


def melted(simulation, n, seed, steps=2000):
    """Melted LJ liquid at T = 1.2: positions, velocities (double, ordered by particle id) and the box."""
    s = simulation.lennard_jones_fluid(n, seed=seed, integrator="andersen", timestep=0.005, rate=2.0)
    s.run(steps)
    order = np.argsort(s.particle.get("id"))
    return s.particle.get("position_ds")[order], s.particle.get("velocity_ds")[order], s.box.length


def product_from(mdsim, simulation, pos, vel, n, seed):
    """A fresh NVE system of the product carrying exactly the given state (ids 0..n-1 in array order)."""
    s = simulation.lennard_jones_fluid(n, seed=seed)
    s.particle.set("position_ds", pos)
    s.particle.set("velocity_ds", vel)
    return s


def by_id(s, name):
    return s.particle.get(name)[np.argsort(s.particle.get("id"))]


def energy(s):
    th = s.thermodynamics()
    return th["internal_energy"], th


def test_trajectory_and_rebuild_schedule_match_the_reference_loop(hx, api, refloop):
    """200 steps from the same state: the same number of list rebuilds (the rebuild criterion is evaluated on
    bit-identical displacements until rounding differences of the forces, <= 1e-7, have grown), positions within
    1e-5 sigma, and the state of the reference loop — loaded back into the product — has the same energy."""
    mdsim, simulation = api
    n = 4000
    pos, vel, L = melted(simulation, n, seed=31)
    ours = product_from(mdsim, simulation, pos, vel, n, seed=32)
    ref = refloop.ReferenceLoop(hx, pos, vel, L)
    u0 = ours.neighbour.updates
    ours.force.apply()
    f_ref0 = ref.state()["force"]
    f0 = by_id(ours, "force")
    mag = np.linalg.norm(f_ref0, axis=1)
    rel = np.linalg.norm(f0 - f_ref0, axis=1) / np.maximum(mag, np.sqrt((mag ** 2).mean()))
    assert rel.max() < 1e-6, rel.max()                 # forces at the common initial state (lists built by both sides)
    steps = 200
    ours.run(steps)
    ms, rebuilds = ref.run(steps)
    assert ours.neighbour.updates - u0 - 1 == rebuilds or ours.neighbour.updates - u0 == rebuilds + 1
    st = ref.state()
    d = by_id(ours, "position_ds") - st["position"]
    d -= np.rint(d / L) * L
    assert np.abs(d).max() < 1e-5, np.abs(d).max()
    dv = by_id(ours, "velocity_ds") - st["velocity"]
    assert np.abs(dv).max() < 1e-4, np.abs(dv).max()
    e_ours, _ = energy(ours)
    meter = product_from(mdsim, simulation, st["position"], st["velocity"], n, seed=33)
    e_ref, _ = energy(meter)
    assert abs(e_ours - e_ref) < 1e-6 * abs(e_ref)
    ref.close()


def test_energy_drift_no_worse_than_the_reference_loop(hx, api, refloop):
    """10^5 NVE steps (dt = 0.001, N = 4000, rho = 0.8442, smooth_r4) with the product and with the reference's own
    kernels from the same melted state; the total energy of either trajectory is measured by the same meter (the
    product's thermodynamics on the state).  Both must pass the reference's validation gate
    (thermodynamics.cpp:150-161) and the product's worst relative deviation must not exceed the reference's (a 25 %
    allowance covers the statistical scatter of this measure between two chaotic trajectories)."""
    mdsim, simulation = api
    n, steps, samples = 4000, 100000, 20
    pos, vel, L = melted(simulation, n, seed=41)
    ours = product_from(mdsim, simulation, pos, vel, n, seed=42)
    meter = product_from(mdsim, simulation, pos, vel, n, seed=43)
    ref = refloop.ReferenceLoop(hx, pos, vel, L)
    e0, _ = energy(ours)
    worst_ours = worst_ref = 0.0
    for _ in range(samples):
        ours.run(steps // samples)
        ref.run(steps // samples)
        e, th = energy(ours)
        worst_ours = max(worst_ours, abs(e - e0) / abs(e0))
        st = ref.state()
        meter.particle.set("position_ds", st["position"])
        meter.particle.set("velocity_ds", st["velocity"])
        e_ref, _ = energy(meter)
        worst_ref = max(worst_ref, abs(e_ref - e0) / abs(e0))
    gate = 3e-5 + steps * 3e-13
    print(f"energy drift over {steps} steps: product {worst_ours:.3e}, reference kernels {worst_ref:.3e}, gate {gate:.3e}")
    assert worst_ref < gate, worst_ref
    assert worst_ours < gate, worst_ours
    assert worst_ours <= 1.25 * worst_ref + 1e-7, (worst_ours, worst_ref)
    ref.close()


def test_thermodynamic_sums_against_reference_kernels(hx, cuda, refloop):
    """hx_thermodynamics vs the reference's accumulators kinetic_energy, velocity_of_centre_of_mass, potential_energy and
    virial (observables/gpu/thermodynamics_kernel.cu:35-97) reduced by its reduction kernel in dsfloat."""
    from test_gpu_kernels import Dev
    dev = Dev(cuda)
    lib = refloop.load()
    lib.refgpu_thermodynamics.restype = C.c_int
    n, A = 100000, 100352
    rng = np.random.default_rng(5)
    vel = np.zeros((A, 4), dtype=np.float32)
    vel[:n, :3] = rng.normal(scale=1.1, size=(n, 3))
    vel[:, 3] = 1.0
    vel[:n, 3] = rng.choice(np.array([1.0, 0.7, 2.3], dtype=np.float32), n)
    en = np.zeros(A, dtype=np.float32)
    en[:n] = rng.normal(loc=-6.0, scale=1.0, size=n)
    st = np.zeros(6 * A, dtype=np.float32)
    st.reshape(6, A)[:, :n] = rng.normal(loc=-3.0, scale=2.0, size=(6, n))
    ref = np.zeros(7)
    assert lib.refgpu_thermodynamics(support.ptr(vel), support.ptr(en), support.ptr(st), C.c_size_t(A), C.c_size_t(n),
                                     support.ptr(ref)) == 0
    d_v, d_e, d_s = dev.up(vel), dev.up(en), dev.up(st)
    d_out = dev.zeros(hx.THERMO_DOUBLES, np.float64)
    hx.check(hx.lib.hx_thermodynamics(dev.p(d_v), dev.p(d_e), dev.p(d_s), A, n, dev.p(d_out), dev.stream))
    out = dev.down(d_out)
    # Same per-particle terms.  The reference accumulates in double-single and its result carries up to float-epsilon
    # relative error in some components (measured: 7e-8 in sum m v_y, 4e-10 in sum m v_x); the product sums in double.
    for k in range(7):
        assert abs(out[k] - ref[k]) <= 2e-7 * max(abs(ref[k]), np.sqrt(n)), (k, out[k], ref[k])
    # the product against the exact sums of the same terms (float inner product for the kinetic energy: its rounding
    # is the only inexact step, hence the looser bound on component 0)
    v64, m64 = vel[:n, :3].astype(np.float64), vel[:n, 3].astype(np.float64)
    exact = np.array([(m64 * (v64 ** 2).sum(axis=1)).sum(), *(m64[:, None] * v64).sum(axis=0), m64.sum(),
                      en[:n].astype(np.float64).sum(), st.reshape(6, A)[:3, :n].astype(np.float64).sum()])
    scale = np.array([(m64 * (v64 ** 2).sum(axis=1)).sum(), *np.abs(m64[:, None] * v64).sum(axis=0), m64.sum(),
                      np.abs(en[:n]).astype(np.float64).sum(), np.abs(st.reshape(6, A)[:3, :n]).astype(np.float64).sum()])
    tol = np.array([1e-7, 1e-13, 1e-13, 1e-13, 1e-13, 1e-13, 1e-13])
    assert np.all(np.abs(out[:7] - exact) <= tol * scale), (out[:7] - exact) / scale
