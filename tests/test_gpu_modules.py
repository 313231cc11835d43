"""Module-level parity and physics tests through the public API (halmd_b200.mdsim / observables), i.e. through the
C++ host layer and the C ABI.  They read like the reference's integration tests (test/integration/validation) and
module unit tests."""
import ctypes as C

import numpy as np
import pytest

import support

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(cuda):
    import halmd_b200
    from halmd_b200 import simulation
    return halmd_b200, simulation


def state_bytes(p):
    return [p.get(k).tobytes() for k in ("position_ds", "velocity_ds", "image", "id", "reverse_id")]


def test_fused_verlet_tail_equals_separate_kernels(api):
    """The velocity-Verlet update fused into the tail of the shell force kernel (hx_pair_force_shells_verlet: one launch per
    step, double-buffered positions, two displacement slots) leaves exactly the state of the separate k_verlet +
    k_force_shell launches — one species and the Kob-Andersen table, across list updates and halted chunks, for runs
    split at arbitrary points."""
    halmd_b200, simulation = api
    from halmd_b200 import mdsim
    a = simulation.lennard_jones_fluid(4000, seed=9)
    b = simulation.lennard_jones_fluid(4000, seed=9)
    a.core.set_fused_verlet(True)
    b.core.set_fused_verlet(False)
    for steps in (1, 2, 3, 250, 7, 137):
        a.run(steps)
        b.run(steps)
        assert state_bytes(a.particle) == state_bytes(b.particle), steps
    assert a.neighbour.updates == b.neighbour.updates and a.neighbour.updates >= 5
    assert a.core.stats()["launches"] < 0.7 * b.core.stats()["launches"]
    assert np.array_equal(a.particle.get("force"), b.particle.get("force"))
    ta, tb = a.thermodynamics(), b.thermodynamics()
    assert ta == tb
    # mixture (species table in shared memory), NVE after a short thermostatted start
    systems = []
    for fused in (True, False):
        s = simulation.kob_andersen_mixture(8000, seed=4)
        s.run(100)
        integ = mdsim.integrators.verlet(box=s.box, particle=s.particle, timestep=0.002)
        s.integrator, s.core = integ, mdsim.core.get()
        s.core.set_fused_verlet(fused)
        s.core.set_chunk(2, 17)
        s.run(300)
        systems.append(s)
    assert state_bytes(systems[0].particle) == state_bytes(systems[1].particle)
    assert systems[0].neighbour.updates == systems[1].neighbour.updates


def test_chunked_run_equals_single_steps(api):
    """sampler:run with the device-resident rebuild decision and fused kernels reproduces, bit for bit, the state of
    the reference call order (integrate, displacement check on the host, rebuild, force, finalize per step)."""
    _, simulation = api
    a = simulation.lennard_jones_fluid(4000, seed=5)
    b = simulation.lennard_jones_fluid(4000, seed=5)
    steps = 300
    for _ in range(steps):
        a.core.mdstep()
    b.core.set_chunk(2, 64)
    b.run(steps)
    assert a.neighbour.updates == b.neighbour.updates and a.neighbour.updates >= 3
    assert state_bytes(a.particle) == state_bytes(b.particle)
    # different chunking, same result; also a run split into pieces
    c = simulation.lennard_jones_fluid(4000, seed=5)
    c.core.set_chunk(1, 3)
    c.run(100)
    c.run(1)
    c.run(199)
    assert state_bytes(a.particle) == state_bytes(c.particle)
    assert a.core.step == b.core.step == c.core.step == steps


def test_andersen_chunked_equals_single_steps(api):
    _, simulation = api
    a = simulation.lennard_jones_fluid(4000, seed=6, integrator="andersen", timestep=0.005, rate=2.0)
    b = simulation.lennard_jones_fluid(4000, seed=6, integrator="andersen", timestep=0.005, rate=2.0)
    for _ in range(120):
        a.core.mdstep()
    b.run(120)
    assert a.neighbour.updates == b.neighbour.updates
    assert state_bytes(a.particle) == state_bytes(b.particle)


def test_hilbert_order_and_ids(api, oracle, hx):
    """After a list update the particles are in Hilbert order (stable), ids are a permutation and reverse_id is its
    inverse (particle.cpp:173-218, sorts/hilbert.cpp:71-117)."""
    _, simulation = api
    s = simulation.lennard_jones_fluid(32000, seed=7)
    s.run(50)
    s.neighbour.check()
    p = s.particle
    n = p.nparticle
    ids, rid = p.get("id"), p.get("reverse_id")
    assert np.array_equal(np.sort(ids), np.arange(n))
    assert np.array_equal(rid[ids], np.arange(n))
    # keys of the positions at the time of the last update are non-decreasing; particles have moved a little since,
    # so re-run an update at the current positions and check the order right after it
    pos = p.get("position")
    fresh = simulation.lennard_jones_fluid(n, seed=7)
    fresh.particle.set("position", pos)
    fresh.neighbour.check()
    pos2 = fresh.particle.get("position")
    L = fresh.box.length
    pk, _, A = support.pack_state(pos2.astype(np.float64))
    keys = np.zeros(A, dtype=np.uint32)
    depth = oracle.oracle_hilbert_depth(support.d3(L))
    oracle.oracle_hilbert_keys(support.ptr(pk), C.c_size_t(A), support.f3(L), depth, support.ptr(keys))
    assert np.all(np.diff(keys[:n].astype(np.int64)) >= 0)
    # stable: equal keys keep their previous relative order = ascending previous slot = ascending id here
    fid = fresh.particle.get("id")
    same = np.diff(keys[:n].astype(np.int64)) == 0
    assert np.all(np.diff(fid.astype(np.int64))[same] > 0)
    # permutation equals the oracle's stable sort of the oracle's keys
    pk0, _, _ = support.pack_state(pos.astype(np.float64))
    k0 = np.zeros(A, dtype=np.uint32)
    oracle.oracle_hilbert_keys(support.ptr(pk0), C.c_size_t(A), support.f3(L), depth, support.ptr(k0))
    kk, vv = k0[:n].copy(), np.arange(n, dtype=np.uint32)
    oracle.oracle_sort_pairs_stable(support.ptr(kk), support.ptr(vv), C.c_size_t(n))
    assert np.array_equal(fid, vv)


def test_module_neighbour_lists_match_oracle(api, oracle):
    """neighbour lists built by the module layer (after sorting) equal the oracle's on the same memory order."""
    _, simulation = api
    for layout in ("row", "column"):
        s = simulation.lennard_jones_fluid(4000, seed=8, layout=layout)
        s.run(37)
        pos = s.particle.get("position")
        fresh = simulation.lennard_jones_fluid(4000, seed=8, layout=layout, disable_sorting=True)
        fresh.particle.set("position", pos)
        lists, cnt = fresh.neighbour.lists()
        n, L = 4000, fresh.box.length
        pk, _, A = support.pack_state(pos.astype(np.float64))
        ncell = fresh.neighbour_ncell if hasattr(fresh, "neighbour_ncell") else None
        nc, cl, cs = (C.c_uint * 3)(), (C.c_float * 3)(), C.c_uint()
        oracle.oracle_binning_plan(support.d3(L), C.c_size_t(n), C.c_float(2.5), C.c_float(0.5), C.c_double(0.5), nc, cl, C.byref(cs))
        ncells = int(np.prod(list(nc)))
        g_cell = np.zeros(ncells * cs.value, dtype=np.uint32)
        ccnt = np.zeros(ncells, dtype=np.uint32)
        oracle.oracle_binning_update(support.ptr(pk), C.c_size_t(n), cl, nc, cs.value, support.ptr(g_cell), support.ptr(ccnt))
        size = fresh.neighbour.size
        assert size == 239
        rr = np.array([np.float32(9.0)], dtype=np.float32)
        o_list = np.zeros(n * 256, dtype=np.uint32)
        o_cnt = np.zeros(n, dtype=np.uint32)
        oracle.oracle_nlist_build(support.ptr(g_cell), nc, cs.value, support.ptr(pk), C.c_size_t(n), support.farr(rr), 1,
                                  support.f3(L), support.ptr(o_list), size, C.c_size_t(256), 1, support.ptr(o_cnt))
        assert np.array_equal(cnt, o_cnt)
        rows = o_list.reshape(n, 256)
        for i in range(n):
            assert np.array_equal(lists[i], rows[i, :o_cnt[i]])


def test_thermodynamics_against_host_backend(api, oracle):
    """E_kin, E_pot, virial, pressure of the module layer vs the double-precision host-backend port on the same
    configuration (identical positions / velocities)."""
    _, simulation = api
    n = 4000
    s = simulation.lennard_jones_fluid(n, seed=9)
    s.run(200)
    th = s.thermodynamics()
    pos, vel = s.particle.get("position_ds"), s.particle.get("velocity_ds")
    L = s.box.length
    sim = C.c_void_p(oracle.host_sim_create(C.c_size_t(n), support.d3(L), 1, support.darr([1.0]), support.darr([1.0]),
                                            support.darr([2.5]), 2, C.c_double(0.005), C.c_double(0.5), C.c_double(0.001), 1))
    pos, vel = np.ascontiguousarray(pos), np.ascontiguousarray(vel)
    oracle.host_sim_set(sim, support.ptr(pos), support.ptr(vel), None, None)
    out = np.zeros(6)
    oracle.host_sim_thermodynamics(sim, support.ptr(out))
    oracle.host_sim_destroy(sim)
    assert abs(th["kinetic_energy"] - out[0]) < 1e-6 * abs(out[0])
    assert abs(th["potential_energy"] - out[1]) < 2e-6 * abs(out[1])
    assert abs(th["virial"] - out[2]) < 5e-5 * max(abs(out[2]), 1.0)
    assert abs(th["temperature"] - 2 * out[0] / 3) < 1e-6
    rho = n / np.prod(L)
    assert abs(th["pressure"] - rho * (2 * out[0] / 3 + out[2] / 3)) < 1e-4
    assert np.abs(th["center_of_mass_velocity"]).max() < 1e-6


def test_trajectory_tracks_host_backend(api, oracle):
    """Short-time agreement of the single/double-single GPU path with the double host backend from identical initial
    conditions (chaos amplifies rounding differences exponentially; 200 steps stay well correlated)."""
    _, simulation = api
    n = 4000
    s = simulation.lennard_jones_fluid(n, seed=10)
    s.run(500)                                     # leave the lattice
    p = s.particle
    ids = p.get("id")
    pos0, vel0 = p.get("position_ds"), p.get("velocity_ds")
    L = s.box.length
    sim = C.c_void_p(oracle.host_sim_create(C.c_size_t(n), support.d3(L), 1, support.darr([1.0]), support.darr([1.0]),
                                            support.darr([2.5]), 2, C.c_double(0.005), C.c_double(0.5), C.c_double(0.001), 1))
    a, b = np.ascontiguousarray(pos0), np.ascontiguousarray(vel0)
    oracle.host_sim_set(sim, support.ptr(a), support.ptr(b), None, None)
    steps = 200
    s.run(steps)
    oracle.host_sim_run(sim, C.c_size_t(steps))
    hpos = np.zeros((n, 3))
    oracle.host_sim_get(sim, support.ptr(hpos), None, None, None)
    oracle.host_sim_destroy(sim)
    ids1 = p.get("id")
    gpos = p.get("position_ds")
    # map both to id order
    g = np.zeros((n, 3)); g[ids1] = gpos
    h = np.zeros((n, 3)); h[ids] = hpos
    d = g - h
    d -= np.rint(d / L) * L
    assert np.abs(d).max() < 5e-4
    assert np.sqrt((d ** 2).mean()) < 2e-5


def test_energy_conservation_nve(api):
    """test/integration/validation/thermodynamics.cpp:150-161 — drift gate max|dE|/|E| < 3e-5 + steps * 3e-13 for the
    smooth cutoff with double-single integration; here 10^5 steps, N = 4000, rho = 0.8442."""
    _, simulation = api
    s = simulation.lennard_jones_fluid(4000, seed=11, integrator="andersen", timestep=0.005, rate=2.0)
    s.run(2000)                                    # melt + thermalise at T = 1.2
    from halmd_b200 import mdsim
    mdsim.integrators.verlet(box=s.box, particle=s.particle, timestep=0.001)
    s.core = mdsim.core.get()
    th0 = s.core.thermodynamics()
    e0 = th0["internal_energy"]
    worst = 0.0
    steps = 100000
    for _ in range(20):
        s.core.run(steps // 20)
        th = s.core.thermodynamics()
        worst = max(worst, abs(th["internal_energy"] - e0) / abs(e0))
    assert worst < 3e-5 + steps * 3e-13, worst
    assert 1.0 < th["temperature"] < 1.4
    # the Andersen melt does not conserve momentum; the NVE run must (verlet.cpp: pairwise forces only)
    assert np.abs(np.subtract(th["center_of_mass_velocity"], th0["center_of_mass_velocity"])).max() < 1e-5


def test_andersen_thermostat_reaches_temperature(api):
    """verlet_nvt_andersen.cpp:103-174 (statistical): temperature and v_cm under the thermostat."""
    _, simulation = api
    s = simulation.lennard_jones_fluid(32000, seed=12, temperature=0.3, integrator="andersen", timestep=0.002, rate=5.0)
    s.integrator.set_temperature(1.5)
    s.run(4000)
    temps = []
    for _ in range(10):
        s.run(100)
        temps.append(s.thermodynamics()["temperature"])
    assert abs(np.mean(temps) - 1.5) < 0.03


def test_hoover_thermostat(api):
    """test/unit/mdsim/integrators/verlet_nvt_hoover.cpp:134-262 with a shorter run: N = 1500 WCA particles at
    rho = 0.1, T 3 -> 1, f_res = 5, dt = 0.002.  The Hamiltonian extended by the chain terms is conserved (the
    reference's GPU gate max(6e-5 here 1e-4 for the mixed accumulators, steps * 1e-12)), the centre of mass stays at
    rest, mean temperature and its fluctuations (kinetic specific heat 3/2) follow the canonical ensemble."""
    _, simulation = api
    n, temp, dt, fres = 1500, 1.0, 0.002, 5.0
    s = simulation.lennard_jones_fluid(n, density=0.1, temperature=temp, start_temperature=3.0, cutoff=2 ** (1 / 6),
                                       truncation="shifted", timestep=dt, skin=0.5, seed=21, integrator="hoover",
                                       resonance_frequency=fres)
    assert s.integrator.mass == pytest.approx([3 * n * temp / (2 * np.pi * fres) ** 2, temp / (2 * np.pi * fres) ** 2])
    steps = int(np.ceil(100 / dt))
    skip = int(np.ceil(20 / dt))
    period = int(round(3 / (fres * dt)))
    s.run(steps // 20)
    th = s.thermodynamics()
    h0 = th["internal_energy"] + s.integrator.internal_energy
    temps, vcm, worst = [], [], 0.0
    for i in range(0, steps, period):
        s.run(period)
        th = s.thermodynamics()
        if i >= skip:
            temps.append(th["temperature"])
        vcm.append(th["center_of_mass_velocity"])
        worst = max(worst, abs(th["internal_energy"] + s.integrator.internal_energy - h0))
    assert worst / abs(h0) < 1e-4, (worst, h0)
    assert np.abs(np.mean(vcm, axis=0)).max() < 1e-5
    c = len(temps)
    assert abs(np.mean(temps) / temp - 1) < 4.5 * np.sqrt(2.0 / (3 * n * (c - 1))) / temp
    cv = 1.5 ** 2 * n * np.var(temps, ddof=1)
    assert abs(cv / 1.5 - 1) < 4.5 * np.sqrt(1.5 * (3 + 6.0 / n) / c) / 1.5
    xi, v_xi = s.integrator.position, s.integrator.velocity
    assert len(xi) == 2 and len(v_xi) == 2 and xi[0] != 0
    # set_temperature rescales the chain masses (verlet_nvt_hoover.cpp:79-95); set_mass overrides them
    s.integrator.set_temperature(2.0)
    assert s.integrator.mass[1] == pytest.approx(2.0 / (2 * np.pi * fres) ** 2)
    s.integrator.set_mass([3.0, 0.5])
    assert s.integrator.mass == [3.0, 0.5]


def test_kob_andersen_mixture(api):
    """BASELINE config 3 (scaled down): binary mixture with Andersen thermostat, then NVE energy conservation."""
    _, simulation = api
    s = simulation.kob_andersen_mixture(32000, seed=13)
    assert s.neighbour.size >= 64
    s.integrator.set_timestep(0.004)
    s.run(1500)
    s.integrator.set_timestep(0.001)
    s.run(500)
    th = s.thermodynamics()
    assert abs(th["temperature"] - 0.7) < 0.05
    from halmd_b200 import mdsim
    mdsim.integrators.verlet(box=s.box, particle=s.particle, timestep=0.001)
    core = mdsim.core.get()
    e0 = core.thermodynamics()["internal_energy"]
    core.run(5000)
    e1 = core.thermodynamics()["internal_energy"]
    assert abs(e1 - e0) / abs(e0) < 3e-5
    sp = s.particle.get("species")
    assert sp.sum() == 32000 - int(0.8 * 32000)


def test_kob_andersen_config3_full_size(api):
    """BASELINE config 3 at full size: 10^6 particles, 80:20 mixture, Andersen thermostat, then NVE.  Size-independent
    properties: the thermostat holds T = 0.7; ids and the species census survive the Hilbert re-orderings; every pair
    inside its species pair's cutoff at the current positions is in the (displacement-valid) neighbour list of sampled
    particles; NVE conserves energy and momentum."""
    _, simulation = api
    from halmd_b200 import mdsim
    n = 1000000
    s = simulation.kob_andersen_mixture(n, seed=2)
    s.integrator.set_timestep(0.005)
    s.run(300)
    s.integrator.set_timestep(0.001)
    s.run(100)
    assert abs(s.thermodynamics()["temperature"] - 0.7) < 0.05
    p = s.particle
    ids = p.get("id").astype(np.int64)
    assert np.array_equal(np.sort(ids), np.arange(n))
    species = p.get("species").astype(np.int64)
    assert np.array_equal(species, (ids >= int(0.8 * n)).astype(np.int64))        # the species travels with the id
    # completeness of the lists at the current positions (no update pending: the criterion is checked by lists())
    lists, cnt = s.neighbour.lists()
    pos = p.get("position").astype(np.float64)
    L = np.asarray(s.box.length, dtype=np.float64)
    r_cut = s.potential.r_cut.astype(np.float64)
    rng = np.random.default_rng(4)
    for i in rng.choice(n, size=48, replace=False):
        d = pos - pos[i]
        d -= np.rint(d / L) * L
        rr = (d ** 2).sum(axis=1)
        inside = np.nonzero(rr < r_cut[species[i], species] ** 2)[0]
        inside = inside[inside != i]
        assert np.all(np.isin(inside, lists[i])), i
        assert 40 < cnt[i] < s.neighbour.size
    integ = mdsim.integrators.verlet(box=s.box, particle=p, timestep=0.001)
    core = mdsim.core.get()
    th0 = core.thermodynamics()
    core.run(400)
    th1 = core.thermodynamics()
    assert abs(th1["internal_energy"] - th0["internal_energy"]) < 3e-5 * abs(th0["internal_energy"])
    assert np.allclose(th1["center_of_mass_velocity"], th0["center_of_mass_velocity"], rtol=0, atol=1e-6)
    assert s.neighbour.updates >= 20


blocking_reference = support.blocking_reference


@pytest.mark.parametrize("case", [(0.12, 4, 5, 2, 3, 37, 240), (0.9, 10, 10, 3, 10, 0, 1210), (0.4, 3, 3, 2, 50, 11, 700)])
def test_blocking_scheme_pairs_against_reference_classes(api, ref, case):
    """sampler:run with a blocking scheme (hxm_core_run_sampling + hxm_blocking): the sequence of correlated sample pairs
    (level, step of the first sample of the block, step of the later sample) equals that of the reference's own clock,
    blocking_scheme, sample-buffer and correlation classes compiled in place (oracle/ref_dynamics.cpp:
    blocking_scheme.cpp:34-200, samples/blocking_scheme.hpp, correlation.hpp:208-224) driven like sampler.cpp:63-80."""
    halmd_b200, simulation = api
    from halmd_b200 import observables
    max_lag, every, size, shift, sep, start, steps = case
    s = simulation.lennard_jones_fluid(1500, seed=3)
    core = halmd_b200.mdsim.core.get()
    s.run(start)
    dt = s.integrator.timestep
    step0 = core.step
    bs = observables.dynamics.blocking_scheme(max_lag=max_lag, every=every, size=size, shift=shift, separation=sep)
    bs.trace()
    observables.sampler.run(steps, blocking_scheme=bs)
    before = bs.trace()
    bs.finalise()
    after = bs.trace(enable=False)
    for fin, mine in ((False, before), (True, after)):
        iv, ev, cnt, _, time = support.ref_blocking_schedule(ref, dt, max_lag, every * dt, size, shift, sep, step0, steps, finalise=fin)
        assert bs.interval == iv
        assert np.array_equal(mine.astype(np.int64), ev), (case, fin)
        assert np.array_equal(bs.time, time)
    msd = bs.mean_square_displacement()
    assert np.array_equal(msd["count"], cnt * 1500)


def test_blocking_scheme_msd_and_isf(api):
    """Blocking-scheme schedule, MSD and ISF accumulation vs an independent numpy evaluation on positions sampled
    through the API at the same steps (the reference has no test for these modules)."""
    halmd_b200, simulation = api
    from halmd_b200 import observables
    n = 4000
    s = simulation.lennard_jones_fluid(n, seed=14)
    s.run(300)
    dt, steps = 0.001, 240
    wv = observables.utility.wavevector(box=s.box, wavenumber=[2 * np.pi / s.box.length[0] * 2, 7.2], tolerance=0.02, max_count=5)
    bs = observables.dynamics.blocking_scheme(max_lag=0.12, every=4, size=5, separation=3, wavevector=wv)
    interval, events = blocking_reference(steps, dt, 0.12, 4 * dt, 5, 2, 3)
    assert bs.interval == interval
    # sample trajectory step by step through the API to have an independent record
    core = halmd_b200.mdsim.core.get()
    step0 = core.step
    record = {}
    L = s.box.length

    def snapshot():
        p = s.particle
        ids = p.get("id")
        r = p.get("position").astype(np.float32)
        img = p.get("image").astype(np.float32)
        unfolded = np.zeros((n, 3), dtype=np.float32)
        unfolded[ids] = (img.astype(np.float64) * L.astype(np.float32).astype(np.float64) + r.astype(np.float64)).astype(np.float32)
        wrapped = np.zeros((n, 3), dtype=np.float32)
        wrapped[ids] = r
        vel = np.zeros((n, 3), dtype=np.float32)
        vel[ids] = p.get("velocity").astype(np.float32)
        return unfolded, wrapped, vel

    due = sorted({st for _, blk in events for st in blk})
    for k in range(steps + 1):
        if k in due:
            record[k] = snapshot()
            bs.sample()
        if k < steps:
            core.run(1)
    bs.finalise()
    assert core.step == step0 + steps
    msd = bs.mean_square_displacement()
    isf = bs.intermediate_scattering_function()
    nlev = len(interval)
    exp_cnt = np.zeros((nlev, 5))
    exp_sum = np.zeros((nlev, 5))
    exp_isf = np.zeros((nlev, 5, len(wv.shell)))
    exp_isf_cnt = np.zeros((nlev, 5, len(wv.shell)))
    q = wv.value.astype(np.float32)
    rho = {}
    exp_mqd = np.zeros((nlev, 5))
    exp_vacf = np.zeros((nlev, 5))
    for st, (unf, wr, _) in record.items():
        qr = wr.astype(np.float64) @ q.astype(np.float64).T
        rho[st] = np.stack([np.cos(qr).sum(axis=0), -np.sin(qr).sum(axis=0)], axis=1)
    for lev, blk in events:
        first = record[blk[0]][0].astype(np.float64)
        for lag, st in enumerate(blk):
            d = record[st][0].astype(np.float64) - first
            exp_cnt[lev, lag] += n
            exp_sum[lev, lag] += (d ** 2).sum()
            exp_mqd[lev, lag] += ((d ** 2).sum(axis=1) ** 2).sum()
            exp_vacf[lev, lag] += (record[blk[0]][2].astype(np.float64) * record[st][2].astype(np.float64)).sum()
            for si, (a, b) in enumerate(wv.shell):
                re = (rho[blk[0]][a:b, 0] * rho[st][a:b, 0] + rho[blk[0]][a:b, 1] * rho[st][a:b, 1]) / n
                exp_isf[lev, lag, si] += re.sum()
                exp_isf_cnt[lev, lag, si] += b - a
    assert np.array_equal(msd["count"], exp_cnt)
    nz = exp_cnt > 0
    assert np.allclose(msd["mean"][nz], (exp_sum / np.maximum(exp_cnt, 1))[nz], rtol=1e-5, atol=1e-9)
    assert np.all(msd["mean"][:, 0][exp_cnt[:, 0] > 0] == 0)
    mqd, vacf = bs.mean_quartic_displacement(), bs.velocity_autocorrelation()
    assert np.array_equal(mqd["count"], exp_cnt) and np.array_equal(vacf["count"], exp_cnt)
    assert np.allclose(mqd["mean"][nz], (exp_mqd / np.maximum(exp_cnt, 1))[nz], rtol=1e-5, atol=1e-12)
    assert np.allclose(vacf["mean"][nz], (exp_vacf / np.maximum(exp_cnt, 1))[nz], rtol=1e-5, atol=1e-7)
    assert np.all(vacf["mean"][:, 0][exp_cnt[:, 0] > 0] > 1.0)            # <v^2> = 3 T at lag 0 (T ~ 0.7 after melting)
    assert np.array_equal(isf["count"], exp_isf_cnt)
    nz = exp_isf_cnt > 0
    assert np.allclose(isf["mean"][nz], (exp_isf / np.maximum(exp_isf_cnt, 1))[nz], rtol=2e-4, atol=2e-4)
    # F(q, 0) = S(q) > 0 and decays with lag at the structure-factor peak
    assert np.all(isf["mean"][0, 0] > 0)


def test_run_sampling_matches_manual_sampling(api):
    """sampler:run(steps) with the blocking scheme connected equals sampling by hand at every step."""
    halmd_b200, simulation = api
    from halmd_b200 import observables
    res = []
    for mode in ("auto", "manual"):
        s = simulation.lennard_jones_fluid(4000, seed=15)
        s.run(100)
        bs = observables.dynamics.blocking_scheme(max_lag=0.2, every=5, size=4, separation=2)
        core = halmd_b200.mdsim.core.get()
        if mode == "auto":
            observables.sampler.run(150, blocking_scheme=bs)
        else:
            for k in range(151):
                bs.sample()
                if k < 150:
                    core.run(1)
        bs.finalise()
        res.append(bs.mean_square_displacement())
    assert np.array_equal(res[0]["count"], res[1]["count"])
    assert np.array_equal(res[0]["mean"], res[1]["mean"])


def slit_pore(n=4000, D=12.0, seed=5, wall=True, timestep=0.002):
    """examples/definitions/slit_pore.lua:51-67 scaled down: LJ fluid between two planar walls with normals -+x at
    x = +-(D + 1) / 2, periodic in y and z."""
    from halmd_b200 import mdsim
    rng = np.random.default_rng(seed)
    Lx = D + 10.0
    Ly = (n / 0.6 / D) ** 0.5
    box = mdsim.box(length=[Lx, Ly, Ly])
    p = mdsim.particle(dimension=3, particles=n, species=1)
    ext = np.array([D - 2.0, Ly, Ly])
    a = (np.prod(ext) / n) ** (1 / 3)                      # grid with near-equal spacing in the slab |x| < D/2 - 1
    mx = max(int(ext[0] / a), 1)
    my = int(np.ceil(np.sqrt(n / mx)))
    g = np.stack(np.meshgrid(np.arange(mx), np.arange(my), np.arange(my), indexing="ij"), -1).reshape(-1, 3)
    assert len(g) >= n
    g = g[rng.permutation(len(g))[:n]].astype(np.float64)
    p.set("position", ((g + 0.5) / np.array([mx, my, my]) * ext - 0.5 * ext).astype(np.float32))
    v = rng.normal(scale=0.9, size=(n, 3))
    p.set("velocity_ds", v - v.mean(axis=0))
    pot = mdsim.potentials.pair.lennard_jones().truncate("smooth_r4", cutoff=2.5, h=0.005)
    force = mdsim.forces.pair(box=box, particle=p, potential=pot)
    integ = mdsim.integrators.verlet_nvt_andersen(box=box, particle=p, timestep=timestep, temperature=0.9, rate=5.0, seed=3)
    core = mdsim.core.get()
    e_pair = core.thermodynamics()["potential_energy"]
    if wall:
        w = mdsim.potentials.external.planar_wall(offset=[-(D + 1) / 2] * 2, surface_normal=[[-1, 0, 0], [1, 0, 0]], epsilon=1.0,
                                                 sigma=1.0, wetting=1.0, cutoff=2.5, smoothing=0.005, species=1)
        mdsim.forces.external(box=box, particle=p, potential=w)
    return box, p, force, integ, e_pair


def test_slit_pore_external_wall(api):
    """mdsim.forces.external + potentials.external.planar_wall on top of the pair force: the wall term enters the
    potential energy, the walls confine the fluid during an Andersen melt, NVE conserves the total energy with the
    extra force kernel in the chunked step loop, and chunked stepping stays bit-identical to single steps."""
    from halmd_b200 import mdsim
    D = 12.0
    box, p, force, integ, e_pair = slit_pore(D=D)
    core = mdsim.core.get()
    e_both = core.thermodynamics()["potential_energy"]
    assert e_both != e_pair and abs(e_both - e_pair) < 5.0              # wall energy per particle is O(0.1)
    core.run(1500)
    assert np.abs(p.get("position")[:, 0]).max() < (D + 1) / 2          # nobody crossed a wall plane
    mdsim.integrators.verlet(box=box, particle=p, timestep=0.001)
    core = mdsim.core.get()
    e0 = core.thermodynamics()["internal_energy"]
    core.run(2000)
    e1 = core.thermodynamics()["internal_energy"]
    assert abs(e1 - e0) < 2e-4 * abs(e0), (e0, e1)
    assert np.abs(p.get("position")[:, 0]).max() < (D + 1) / 2
    # chunked == single steps (Andersen melt, 60 steps)
    states = []
    for single in (False, True):
        box, p, force, integ, _ = slit_pore(D=D, seed=6)
        core = mdsim.core.get()
        if single:
            for _ in range(60):
                core.run(1)
        else:
            core.run(60)
        states.append(state_bytes(p))
    assert states[0] == states[1]


def test_wetting_slab_config5(api):
    """BASELINE config 5 at full size: N = 2,000,000 LJ particles between two planar walls (slit pore), Andersen melt,
    NVE, density modes along the wall normal and the static structure factor parallel to the walls.  Size-independent
    properties: confinement, energy conservation, layering (the fluid orders against the walls: the density mode at
    the layer spacing along x dominates the same wavenumber parallel to the walls), particle conservation."""
    from halmd_b200 import mdsim, observables
    n, D = 2000000, 60.0
    box, p, force, integ, e_pair = slit_pore(n=n, D=D, seed=8, timestep=0.002)
    core = mdsim.core.get()
    core.run(400)                                                          # Andersen melt at T = 0.9
    x = p.get("position")[:, 0]
    assert np.abs(x).max() < (D + 1) / 2
    mdsim.integrators.verlet(box=box, particle=p, timestep=0.001)
    core = mdsim.core.get()
    e0 = core.thermodynamics()["internal_energy"]
    core.run(300)
    e1 = core.thermodynamics()["internal_energy"]
    assert abs(e1 - e0) < 1e-4 * abs(e0), (e0, e1)
    assert np.array_equal(np.sort(p.get("id")), np.arange(n))
    Lx, Ly = box.length[0], box.length[1]
    mx = int(round(Lx / 0.93))                                             # layer spacing ~ 0.93 sigma at the walls
    class wv:                                                              # two hand-picked wavevectors
        value = np.array([[2 * np.pi / Lx * mx, 0, 0], [0, 2 * np.pi / Ly * int(round(Ly / 0.93)), 0]])
        shell = [(0, 1), (1, 2)]
    s = observables.ssf(density_mode=observables.density_mode(particle=p, wavevector=wv), norm=n).sample()
    assert s[0, 0] > 3 * s[1, 0], s


def test_ssf_of_fcc_lattice(api):
    """test/unit/observables/ssf.cpp:100-300: density modes and static structure factor of a perfect fcc lattice —
    rho_q vanishes unless q is a reciprocal lattice vector with (h, k, l) all even or all odd, where |rho_q| = N and
    S(q) = N; tolerance 50 N epsilon on rho_q as in the reference test."""
    from halmd_b200 import mdsim, observables
    from halmd_b200 import simulation
    ncell = np.array([6, 12, 12])
    n = 4 * int(np.prod(ncell))
    a = (4 / 0.3) ** (1 / 3)
    box = mdsim.box(length=list(ncell * a))
    p = mdsim.particle(dimension=3, particles=n, species=1)
    pos = simulation.fcc_positions(n, box.length)
    p.set("position", pos.astype(np.float32))
    q_lat = 2 * np.pi / a
    wavenumber = sorted(set([q_lat / ncell[0], q_lat / ncell[1]] + [np.sqrt(m) * q_lat for m in (1, 2, 3, 4, 5, 6, 8, 9, 10, 12)]))
    wv = observables.utility.wavevector(box=box, wavenumber=wavenumber, tolerance=1e-3, max_count=6)
    dm = observables.density_mode(particle=p, wavevector=wv)
    rho = dm.acquire()
    assert len(rho) == len(wv.value)
    hkl = np.rint(wv.value * a / (2 * np.pi) * 1e6) / 1e6
    integer = np.all(np.abs(hkl - np.rint(hkl)) < 1e-6, axis=1)
    parity = np.rint(hkl).astype(int) % 2
    bragg = integer & ((parity.sum(axis=1) == 0) | (parity.sum(axis=1) == 3))
    tol = 50 * n * 1.1920929e-07
    assert bragg.any() and (~bragg).any()
    assert np.all(np.abs(np.abs(rho[bragg]) - n) < tol)
    assert np.all(np.abs(rho[~bragg]) < tol)
    s = observables.ssf(density_mode=dm, norm=n).sample()
    for k, (lo, hi) in enumerate(wv.shell):
        expect = n * bragg[lo:hi].mean()
        assert abs(s[k, 0] - expect) < 2 * tol + 1e-3 * expect and s[k, 2] == hi - lo


def test_hard_core_potential_through_modules(api):
    """potential:modify({"hard_core", radius = 0.3}):truncate({"shifted", cutoff = 2^(1/6) + 0.3}) — soft spheres around
    a core (lennard_jones.lua.in:120-130): forces equal those of plain WCA particles evaluated at r - r_core and scaled
    by (r - r_core) / r (checked pair by pair on a dilute configuration), and an NVE run conserves energy."""
    halmd_b200, simulation = api
    from halmd_b200 import mdsim
    n, rho = 4000, 0.3
    L = (n / rho) ** (1 / 3)
    rng = np.random.default_rng(31)
    box = mdsim.box(length=[L, L, L])
    particle = mdsim.particle(dimension=3, particles=n, species=1)
    pos = simulation.fcc_positions(n, box.length).astype(np.float32)
    particle.set("position", pos)
    particle.set("velocity_ds", simulation.boltzmann_velocities(n, 1.0, rng))
    core, cut = 0.3, 2 ** (1 / 6) + 0.3
    pot = mdsim.potentials.pair.lennard_jones().modify("hard_core", radius=core)
    pot = pot.truncate("shifted", cutoff=cut)
    assert pot.params.hard_core == 1 and pot.params.r_core[0] == np.float32(core)
    nb = mdsim.neighbour(particle=particle, box=box, r_cut=pot.r_cut, skin=0.5)
    force = mdsim.forces.pair(box=box, particle=particle, potential=pot, neighbour=nb)
    integ = mdsim.integrators.verlet(box=box, particle=particle, timestep=0.002)
    core_ = mdsim.core.get()
    core_.run(500)
    th0 = core_.thermodynamics()
    core_.run(5000)
    th = core_.thermodynamics()
    assert abs(th["internal_energy"] - th0["internal_energy"]) < 2e-4 * abs(th0["internal_energy"]) + 2e-5
    assert th["potential_energy"] > 0          # purely repulsive, shifted to zero at the cutoff
    # brute-force forces in double on the final configuration
    r = particle.get("position").astype(np.float32)
    f = particle.get("force").astype(np.float64)[:, :3]
    sample = rng.choice(n, 200, replace=False)
    Lf = np.float32(L)
    for i in sample:
        # separations as the kernels see them: single-precision difference and minimum image (a pair across the
        # boundary carries the rounding of a difference of magnitude L), everything after that in double
        d = r[i] - r
        d = (d - np.rint(d / Lf) * Lf).astype(np.float64)
        dist = np.sqrt((d ** 2).sum(axis=1))
        m = (dist > 0) & (dist < cut)
        rs = dist[m] - core
        fv = 24 * (2 * rs ** -12 - rs ** -6) / rs / dist[m]
        fi = (fv[:, None] * d[m]).sum(axis=0)
        assert np.abs(fi - f[i]).max() < 2e-5 * max(1.0, np.abs(fi).max())


def test_api_errors(api):
    halmd_b200, simulation = api
    from halmd_b200 import mdsim, HxError
    box = mdsim.box(length=[5.0, 5.0, 5.0])
    p = mdsim.particle(particles=100)
    pot = mdsim.potentials.pair.lennard_jones().truncate(cutoff=2.5)
    assert pot.truncation == "shifted"                       # default truncation (truncate.lua.in:289)
    with pytest.raises(HxError):
        mdsim.neighbour(particle=p, box=box, r_cut=pot.r_cut)     # fewer than 3 cells per dimension
    with pytest.raises(HxError):
        p.get("nonsense")
    with pytest.raises(HxError):
        mdsim.forces.pair(box=box, particle=p, potential=mdsim.potentials.pair.lennard_jones())
    with pytest.raises(ValueError):
        mdsim.particle(particles=10, dimension=2)
