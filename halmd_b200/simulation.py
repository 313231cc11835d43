"""Synthetic inputs and canned system set-ups for the BASELINE configurations (bench.py, smoke, tests).

Only the *inputs* are generated here (numpy): fcc lattice as halmd.mdsim.positions.lattice
(halmd/mdsim/gpu/positions/lattice.cpp:116-177, lattice_primitive.hpp:121-129) and Maxwell-Boltzmann velocities with
zero centre-of-mass velocity as halmd.mdsim.velocities.boltzmann.  All simulation work goes through the module API.
"""
import ctypes as C

import numpy as np

from .api import mdsim, observables


def fcc_lattice(n, L):
    """Lattice constant, unit cells per dimension and number of sites to skip for n particles in a box L
    (halmd/mdsim/gpu/positions/lattice.cpp:116-177)."""
    L = np.asarray(L, dtype=np.float64)
    u = 4
    a = (np.prod(L) / np.ceil(n / u)) ** (1.0 / 3)
    ncell = np.floor(L / a).astype(np.int64)
    while n > u * np.prod(ncell):
        t = L / (ncell + 1)
        a = t.max()
        m = ncell.copy()
        ncell = np.maximum(m, np.floor(L / a).astype(np.int64))
        if np.all(m == ncell):
            ncell += 1
    return a, ncell, int(u * np.prod(ncell))


def fcc_sites(k, n, L, lattice=None):
    """Positions of particles k (array of particle numbers in [0, n)) on the fcc lattice: surplus sites are left
    vacant at regular intervals, then lattice_primitive.hpp:121-129 maps site index -> position."""
    L = np.asarray(L, dtype=np.float64)
    a, ncell, N = lattice or fcc_lattice(n, L)
    k = np.asarray(k, dtype=np.int64)
    if N > n:
        skip = int(np.ceil(N / (N - n)))
        idx = k + k // (skip - 1)             # every skip-th site (skip-1, 2*skip-1, ...) stays vacant
    else:
        idx = k
    cell = idx >> 2
    r = np.empty((len(k), 3), dtype=np.float64)
    r[:, 0] = cell % ncell[0] + (((idx ^ (idx >> 1)) & 1) * 2 + 1) / 4.0
    r[:, 1] = (cell // ncell[0]) % ncell[1] + ((idx & 1) * 2 + 1) / 4.0
    r[:, 2] = cell // (ncell[0] * ncell[1]) + ((idx & 2) + 1) / 4.0
    r *= a
    r += (L - ncell * a) / 2 - L / 2
    return r


def fcc_positions(n, L):
    return fcc_sites(np.arange(n, dtype=np.int64), n, L)


def boltzmann_velocities(n, temperature, rng, mass=None):
    v = rng.normal(scale=np.sqrt(temperature), size=(n, 3))
    if mass is not None:
        v /= np.sqrt(np.asarray(mass))[:, None]
    v -= v.mean(axis=0)
    return v


class System:
    """A wired-up set of modules, as a HALMD script's main() would build them."""

    def __init__(self, box, particle, potential, force, integrator):
        self.box, self.particle, self.potential, self.force, self.integrator = box, particle, potential, force, integrator
        self.core = mdsim.core.get()
        self.neighbour = force.neighbour

    def run(self, steps):
        self.core.run(steps)

    def thermodynamics(self):
        return self.core.thermodynamics()


def lennard_jones_fluid(n, density=0.8442, temperature=1.2, cutoff=2.5, truncation="smooth_r4", h=0.005, timestep=0.001,
                        skin=0.5, seed=1, layout="row", occupancy=None, integrator="verlet", rate=2.0,
                        disable_sorting=False, shuffle=True, resonance_frequency=5.0, start_temperature=None):
    """BASELINE configs 1/2: one-component LJ fluid on an fcc lattice with Maxwell-Boltzmann velocities."""
    L = (n / density) ** (1.0 / 3)
    rng = np.random.default_rng(seed)
    box = mdsim.box(length=[L, L, L])
    particle = mdsim.particle(dimension=3, particles=n, species=1)
    pos = fcc_positions(n, box.length)
    if shuffle:
        pos = pos[rng.permutation(n)]          # particle order carries no spatial information initially
    particle.set("position", pos.astype(np.float32))
    particle.set("velocity_ds", boltzmann_velocities(n, temperature if start_temperature is None else start_temperature, rng))
    potential = mdsim.potentials.pair.lennard_jones()
    potential = potential.truncate(truncation, cutoff=cutoff, h=h) if truncation == "smooth_r4" else \
        potential.truncate(truncation, cutoff=cutoff)
    nb = mdsim.neighbour(particle=particle, box=box, r_cut=potential.r_cut, skin=skin, occupancy=occupancy, layout=layout,
                         disable_sorting=disable_sorting)
    force = mdsim.forces.pair(box=box, particle=particle, potential=potential, neighbour=nb)
    if integrator == "verlet":
        integ = mdsim.integrators.verlet(box=box, particle=particle, timestep=timestep)
    elif integrator == "hoover":
        integ = mdsim.integrators.verlet_nvt_hoover(box=box, particle=particle, timestep=timestep,
                                                    temperature=temperature, resonance_frequency=resonance_frequency)
    else:
        integ = mdsim.integrators.verlet_nvt_andersen(box=box, particle=particle, timestep=timestep,
                                                      temperature=temperature, rate=rate, seed=seed)
    return System(box, particle, potential, force, integ)


def kob_andersen_mixture(n, density=1.2, temperature=0.7, timestep=0.001, skin=0.3, occupancy=0.7, seed=1, rate=2.0,
                         layout="row", integrator="verlet_nvt_andersen"):
    """BASELINE config 3: 80:20 binary mixture, eps = (1, 1.5, 0.5), sigma = (1, 0.8, 0.88), r_c = 2.5 sigma_ab,
    smooth_r4 (examples/benchmarks/kob_andersen)."""
    L = (n / density) ** (1.0 / 3)
    rng = np.random.default_rng(seed)
    box = mdsim.box(length=[L, L, L])
    particle = mdsim.particle(dimension=3, particles=n, species=2)
    pos = fcc_positions(n, box.length)[rng.permutation(n)]
    species = (np.arange(n) >= int(0.8 * n)).astype(np.uint32)     # first 80 % are A, assigned by id
    particle.set("position", pos.astype(np.float32))
    particle.set("species", species)
    particle.set("velocity_ds", boltzmann_velocities(n, temperature, rng))
    potential = mdsim.potentials.pair.lennard_jones(epsilon=[[1.0, 1.5], [1.5, 0.5]], sigma=[[1.0, 0.8], [0.8, 0.88]])
    potential = potential.truncate("smooth_r4", cutoff=2.5, h=0.005)
    nb = mdsim.neighbour(particle=particle, box=box, r_cut=potential.r_cut, skin=skin, occupancy=occupancy, layout=layout)
    force = mdsim.forces.pair(box=box, particle=particle, potential=potential, neighbour=nb)
    if integrator == "verlet":
        integ = mdsim.integrators.verlet(box=box, particle=particle, timestep=timestep)
    else:
        integ = mdsim.integrators.verlet_nvt_andersen(box=box, particle=particle, timestep=timestep,
                                                      temperature=temperature, rate=rate, seed=seed)
    return System(box, particle, potential, force, integ)


def wetting_slab(n, D=None, density=0.6, temperature=0.9, timestep=0.001, seed=5, rate=5.0, integrator="verlet_nvt_andersen"):
    """BASELINE config 5 (examples/definitions/slit_pore.lua:51-67): LJ fluid (smooth_r4, r_c = 2.5) between two planar
    Lennard-Jones walls with normals -+x at x = +-(D + 1) / 2, periodic in y and z; box (D + 10) x L_y x L_y with
    L_y from the slab density.  D defaults to the value that makes the slab about cubic."""
    rng = np.random.default_rng(seed)
    if D is None:
        D = float(np.round((n / density) ** (1.0 / 3)))
    Lx = D + 10.0
    Ly = (n / density / D) ** 0.5
    box = mdsim.box(length=[Lx, Ly, Ly])
    p = mdsim.particle(dimension=3, particles=n, species=1)
    ext = np.array([D - 2.0, Ly, Ly])
    a = (np.prod(ext) / n) ** (1 / 3)                      # grid with near-equal spacing in the slab |x| < D/2 - 1
    mx = max(int(ext[0] / a), 1)
    my = int(np.ceil(np.sqrt(n / mx)))
    g = np.stack(np.meshgrid(np.arange(mx), np.arange(my), np.arange(my), indexing="ij"), -1).reshape(-1, 3)
    g = g[rng.permutation(len(g))[:n]].astype(np.float64)
    p.set("position", ((g + 0.5) / np.array([mx, my, my]) * ext - 0.5 * ext).astype(np.float32))
    v = rng.normal(scale=np.sqrt(temperature), size=(n, 3))
    p.set("velocity_ds", v - v.mean(axis=0))
    pot = mdsim.potentials.pair.lennard_jones().truncate("smooth_r4", cutoff=2.5, h=0.005)
    force = mdsim.forces.pair(box=box, particle=p, potential=pot)
    if integrator == "verlet":
        integ = mdsim.integrators.verlet(box=box, particle=p, timestep=timestep)
    else:
        integ = mdsim.integrators.verlet_nvt_andersen(box=box, particle=p, timestep=timestep, temperature=temperature, rate=rate, seed=seed)
    wall = mdsim.potentials.external.planar_wall(offset=[-(D + 1) / 2] * 2, surface_normal=[[-1, 0, 0], [1, 0, 0]], epsilon=1.0,
                                                 sigma=1.0, wetting=1.0, cutoff=2.5, smoothing=0.005, species=1)
    mdsim.forces.external(box=box, particle=p, potential=wall)
    s = System(box, p, pot, force, integ)
    s.D = D
    return s


def decomposed_lj_fluid(n, rank, nranks, density=0.8442, temperature=1.2, cutoff=2.5, truncation="smooth_r4", h=0.005,
                        timestep=0.001, skin=0.5, seed=1, integrator="verlet", rate=2.0, group=None, chunk=1 << 22,
                        lanes=4):
    """BASELINE config 4: one-component LJ fluid of n particles in total, slab-decomposed over `nranks` processes.
    Every rank generates the lattice and the velocities chunk by chunk (same seeds everywhere) and keeps its slab, so
    no rank ever holds the global state; particle ids are the lattice site numbers."""
    from .decomposition import DecomposedSystem, SlabDecomposition
    L = np.array([(n / density) ** (1.0 / 3)] * 3)
    dec = SlabDecomposition(L, nranks, rank, cutoff + skin)
    lattice = fcc_lattice(n, L)
    pos, vel, ids = [], [], []
    vsum = np.zeros(3)
    for c, c0 in enumerate(range(0, n, chunk)):
        k = np.arange(c0, min(n, c0 + chunk), dtype=np.int64)
        r = fcc_sites(k, n, L, lattice).astype(np.float32)
        v = np.random.default_rng([seed, c]).normal(scale=np.sqrt(temperature), size=(len(k), 3))
        vsum += v.sum(axis=0)
        mine = dec.slab_of(r[:, 0]) == rank
        pos.append(r[mine].astype(np.float64))
        vel.append(v[mine])
        ids.append(k[mine])
    pos, vel, ids = np.concatenate(pos), np.concatenate(vel), np.concatenate(ids)
    vel -= vsum / n
    return DecomposedSystem(L, pos, vel, ids, cutoff=(cutoff,), truncation=truncation, h=h, skin=skin, timestep=timestep,
                            integrator=integrator, temperature=temperature, rate=rate, seed=seed, n_global=n, rank=rank,
                            nranks=nranks, group=group, lanes=lanes)


def smoke_check(hx, oracle, support, torch, np_):
    """One small pass of the hot path on cuda:0 checked against the oracle (used by __graft_entry__.smoke)."""
    n = 4000
    sysm = lennard_jones_fluid(n, seed=3)
    sysm.run(20)                                   # Hilbert order, binning, lists, force, fused Verlet steps
    p = sysm.particle
    pos = p.get("position")
    species = p.get("species")
    lists, cnt = sysm.neighbour.lists()
    # oracle: lists from the same positions
    L = sysm.box.length
    pk, vk, A = support.pack_state(pos.astype(np_.float64), species)
    ncell = sysm.neighbour.binning.ncell if sysm.neighbour.binning else None
    ncell_c, cl, cs = (C.c_uint * 3)(), (C.c_float * 3)(), C.c_uint()
    oracle.oracle_binning_plan(support.d3(L), C.c_size_t(n), C.c_float(2.5), C.c_float(0.5), C.c_double(0.5), ncell_c, cl, C.byref(cs))
    ncells = int(np_.prod(list(ncell_c)))
    g_cell = np_.zeros(ncells * cs.value, dtype=np_.uint32)
    ccnt = np_.zeros(ncells, dtype=np_.uint32)
    assert oracle.oracle_binning_update(support.ptr(pk), C.c_size_t(n), cl, ncell_c, cs.value, support.ptr(g_cell), support.ptr(ccnt)) == 0
    size, stride = sysm.neighbour.size, sysm.neighbour.stride
    rr = np_.array([np_.float32((np_.float32(2.5) + np_.float32(0.5)) ** 2)], dtype=np_.float32)
    o_list = np_.zeros(n * stride, dtype=np_.uint32)
    o_cnt = np_.zeros(n, dtype=np_.uint32)
    # the lists on the device were built at the last update; rebuild both sides at the current positions
    sysm.neighbour  # noqa: B018
    oracle.oracle_nlist_build(support.ptr(g_cell), ncell_c, cs.value, support.ptr(pk), C.c_size_t(n), support.farr(rr), 1,
                              support.f3(L), support.ptr(o_list), size, C.c_size_t(stride), 1, support.ptr(o_cnt))
    # force of the oracle on the device's own lists (lists may stem from an earlier step: still valid Verlet lists)
    raw = np_.full(n * stride, 0xFFFFFFFF, dtype=np_.uint32).reshape(n, stride)
    for i, l in enumerate(lists):
        raw[i, :len(l)] = l
    par = support.OLjParams()
    oracle.oracle_lj_init(C.byref(par), 1, 1, support.farr([1.0]), support.farr([1.0]), 2, support.farr([2.5]), C.c_float(0.005))
    of = np_.zeros((A, 4), dtype=np_.float32)
    oracle.oracle_pair_force(C.byref(par), support.ptr(pk), support.ptr(pk), support.ptr(raw), size, C.c_size_t(stride), 1,
                             C.c_size_t(n), support.f3(L), 1, C.c_float(0.5), support.ptr(of), None, None, C.c_size_t(A))
    sysm.force.apply()
    f = p.get("force")
    d = np_.linalg.norm(f.astype(np_.float64) - of[:n, :3], axis=1)
    mag = np_.linalg.norm(of[:n, :3].astype(np_.float64), axis=1)
    rel = float((d / np_.maximum(mag, np_.sqrt((mag ** 2).mean()))).max())
    assert rel < 3e-6, rel      # IEEE oracle vs the reference's MUFU arithmetic mirrored by the kernels (tests/test_gpu_kernels.py)
    # a fresh list build at the current positions must reproduce the oracle's sets exactly
    fresh = lennard_jones_fluid(n, seed=3, disable_sorting=True, shuffle=False)
    fresh.particle.set("position", pos)
    fl, fc = fresh.neighbour.lists()
    assert np_.array_equal(fc, o_cnt)
    o_rows = o_list.reshape(n, stride)
    for i in range(0, n, 7):
        assert np_.array_equal(fl[i], o_rows[i, :o_cnt[i]])
    th = sysm.thermodynamics()
    return {"n": n, "steps": int(sysm.core.step), "list_updates": int(sysm.neighbour.updates),
            "force_rel_err": rel, "mean_neighbours": float(cnt.mean()), "en_tot": th["internal_energy"],
            "launches": int(hx.lib.hx_launch_count())}
