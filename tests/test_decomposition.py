"""Slab domain decomposition: host logic on CPU over gloo (world sizes 2 and 3), and — on a GPU box —
parity of the decomposed step loop with the single-GPU path (several ranks sharing cuda:0, gloo staging).

The reference cannot run multi-GPU, so the oracle of the decomposed run is the same code on one GPU: identical
neighbour sets per particle id, forces within 1e-6, and, for a single rank, bitwise identical trajectories."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dd_worker.py")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(mode, world, out, extra_env=None, timeout=300):
    port = free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   GLOO_SOCKET_IFNAME="lo", **(extra_env or {}))
        procs.append(subprocess.Popen([sys.executable, WORKER, mode, out], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
    for rank, (p, o) in enumerate(zip(procs, logs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{o[-3000:]}"


# ---- pure host logic ---------------------------------------------------------------------------------------------
def test_slab_ownership_partitions_the_box():
    from halmd_b200.decomposition import SlabDecomposition
    L = [30.0, 10.0, 10.0]
    x = np.linspace(-15, 15, 20001, dtype=np.float32)[:-1]
    for P in (1, 2, 3, 4, 8):
        owners = [SlabDecomposition(L, P, r, 3.0).slab_of(x) for r in range(P)]
        for o in owners[1:]:
            assert np.array_equal(o, owners[0])                 # every rank computes the same owner
        assert owners[0].min() == 0 and owners[0].max() == P - 1
        assert np.all(np.diff(owners[0]) >= 0)                  # slabs are contiguous in x
        counts = np.bincount(owners[0], minlength=P)
        assert counts.max() - counts.min() <= 2                 # and of equal width
        for r in range(P):
            d = SlabDecomposition(L, P, r, 3.0)
            key = d.owner_keys(x)
            assert np.array_equal(key == 0, owners[0] == r)
            if P >= 3:
                assert np.array_equal(key == 1, owners[0] == d.left) and np.array_equal(key == 2, owners[0] == d.right)
            if P > 3:
                assert np.any(key == 3)


def test_cell_layers_cover_the_slab_and_its_halo():
    """SlabDecomposition.cell_layers: every position a rank can hold (its slab widened by the halo, wrapped into the
    box) falls into the layers of the global cell grid the rank processes, with a layer to spare on either side — also
    with the cell index evaluated one ulp off, as the device's approximate reciprocal may (csrc/binning.cu cell_coord)."""
    from halmd_b200.decomposition import SlabDecomposition
    rng = np.random.default_rng(3)
    for Lx, P, halo in ((423.2, 8, 3.0), (120.0, 4, 3.0), (60.0, 2, 3.0), (37.1, 3, 2.8), (30.0, 8, 3.0), (19.0, 2, 3.0), (64.3, 1, 3.0)):
        n = int(Lx // halo)
        ell = Lx / n
        for r in range(P):
            d = SlabDecomposition([Lx, 20.0, 20.0], P, r, halo)
            x0, nx = d.cell_layers(np.float32(ell), n)
            assert 0 <= x0 < n and 1 <= nx <= n
            x = rng.uniform(float(d.lo) - float(d.halo), float(d.hi) + float(d.halo), 200000)
            x = np.concatenate([x, [float(d.lo) - float(d.halo), float(d.hi) + float(d.halo), float(d.lo), float(d.hi)]])
            x = (x - np.rint(x / Lx) * Lx).astype(np.float32)
            for rcp in (np.float32(1) / np.float32(ell), np.nextafter(np.float32(1) / np.float32(ell), np.float32(2)),
                        np.nextafter(np.float32(1) / np.float32(ell), np.float32(0))):
                c = np.floor(x.astype(np.float64) * float(rcp) + n).astype(np.int64) % n
                rel = (c - x0) % n
                assert rel.max() < nx, (Lx, P, r)
                if nx < n:
                    assert rel.min() >= 1 and rel.max() <= nx - 2       # the margin layers stay empty
        if P >= 4 and Lx / P > 6 * halo:
            assert nx < n                                               # a slab processes a fraction of the grid


def test_ghost_selection_folds_the_periodic_image():
    """A coordinate of exactly +L/2 (the integrator's wrap leaves [-L/2, +L/2]) is owned by the slab that starts at -L/2;
    it must be offered to that slab's LEFT neighbour (the last rank) and not to the right one."""
    from halmd_b200.decomposition import SlabDecomposition
    L = [96.0, 10.0, 10.0]
    x = np.array([48.0, -48.0, -47.0, -42.0, -36.5, 47.9], dtype=np.float32)
    d0 = SlabDecomposition(L, 8, 0, 3.0)
    assert np.array_equal(d0.slab_of(x), [0, 0, 0, 0, 0, 7])
    own = x[:5]
    assert np.array_equal(d0.ghost_mask(own, 0), [True, True, True, False, False])        # towards rank 7
    assert np.array_equal(d0.ghost_mask(own, 1), [False, False, False, False, True])       # towards rank 1
    d7 = SlabDecomposition(L, 8, 7, 3.0)
    assert np.array_equal(d7.ghost_mask(x[5:], 1), [True]) and np.array_equal(d7.ghost_mask(x[5:], 0), [False])
    # a particle an ulp below its slab's low face (rounding of lo = x0 + rank * width) still goes to the left neighbour
    d5 = SlabDecomposition([423.2317, 10.0, 10.0], 8, 5, 3.0)
    xb = np.array([np.nextafter(d5.lo, np.float32(-1e9)), d5.lo, np.nextafter(d5.hi, np.float32(-1e9))], dtype=np.float32)
    assert np.array_equal(d5.ghost_mask(xb, 0), [True, True, False]) and np.array_equal(d5.ghost_mask(xb, 1), [False, False, True])
    # two ranks: the single combined list holds every boundary particle once
    d = SlabDecomposition(L, 2, 0, 3.0)
    assert np.array_equal(d.ghost_mask(np.array([48.0, -48.0, -1.0, -24.0], dtype=np.float32), 2), [True, True, True, False])


def test_slab_rejects_thin_slabs():
    from halmd_b200.decomposition import SlabDecomposition
    with pytest.raises(ValueError):
        SlabDecomposition([20.0, 20.0, 20.0], 8, 0, 3.0)         # width 2.5 < r_c + skin
    with pytest.raises(ValueError):
        SlabDecomposition([5.0, 20.0, 20.0], 2, 0, 3.0)
    SlabDecomposition([24.1, 20.0, 20.0], 8, 7, 3.0)


def test_ghost_sides_two_ranks_send_each_particle_once():
    from halmd_b200.decomposition import SlabDecomposition
    d = SlabDecomposition([10.0, 10.0, 10.0], 2, 0, 3.0)         # slab width 5 < 2 * halo: both faces overlap
    x = np.linspace(d.lo, d.hi, 1000, dtype=np.float32)[:-1]
    assert d.ghost_sides() == [(2, "left")]
    assert d.ghost_mask(x, 2).all()
    d3 = SlabDecomposition([30.0, 10.0, 10.0], 3, 1, 3.0)
    assert d3.ghost_sides() == [(0, "left"), (1, "right")]
    x = np.linspace(d3.lo, d3.hi, 1000, dtype=np.float32)[:-1]
    assert 0.29 < d3.ghost_mask(x, 0).mean() < 0.32 and 0.29 < d3.ghost_mask(x, 1).mean() < 0.32
    assert not np.any(d3.ghost_mask(x, 0) & d3.ghost_mask(x, 1))


@pytest.mark.parametrize("world", [2, 3])
def test_migration_and_ghost_protocol_gloo(tmp_path, world):
    """world_size-2/3 gloo run of the exchange protocol: ownership stays a partition, ghost layers are complete."""
    out = str(tmp_path / "proto.npz")
    launch("cpu_protocol", world, out)
    r = np.load(out)
    assert len(r.files) == 4
    for k in r.files:
        n_own, n_ghost, moved = r[k]
        assert n_own > 0 and n_ghost > 0 and moved > 0


# ---- decomposed step loop vs. single GPU ------------------------------------------------------------------------------
def single_gpu_reference(n, steps):
    from halmd_b200 import simulation
    s = simulation.lennard_jones_fluid(n, seed=21)            # same generator calls as dd_worker.gpu_parity
    p = s.particle

    def snap():
        s.force.apply()
        ids = p.get("id").astype(np.int64)
        order = np.argsort(ids)
        lists, cnt = s.neighbour.lists()
        sets = [np.sort(ids[lists[i]]) for i in range(n)]
        state = {"position": p.get("position_ds")[order], "velocity": p.get("velocity_ds")[order],
                 "force": p.get("force")[order], "sets": [sets[i] for i in order]}
        th = s.thermodynamics()               # (re-evaluates the forces with auxiliary variables: read them before)
        state["thermo"] = np.array([th["internal_energy"], th["potential_energy"], th["temperature"], th["pressure"]])
        return state
    a = snap()
    s.run(steps)
    b = snap()
    return a, b, s.neighbour.updates


def compare(ref, r, tag, n, bitwise, bitwise_force):
    sets = np.split(r[f"{tag}_nl_flat"], np.cumsum(r[f"{tag}_nl_count"])[:-1])
    if bitwise:
        assert np.array_equal(ref["position"], r[f"{tag}_position"])
        assert np.array_equal(ref["velocity"], r[f"{tag}_velocity"])
    if bitwise_force:
        assert np.array_equal(ref["force"], r[f"{tag}_force"])
    f, g = ref["force"].astype(np.float64), r[f"{tag}_force"].astype(np.float64)
    mag = np.linalg.norm(f, axis=1)
    # forces nearly cancel on the initial lattice: the scale of the comparison is a typical pair force, not |f_i|
    rel = np.linalg.norm(f - g, axis=1) / np.maximum(np.maximum(mag, np.sqrt((mag ** 2).mean())), 10.0)
    return sets, rel.max()


@pytest.mark.gpu
@pytest.mark.parametrize("world,n,backend", [(1, 4000, "gloo"), (2, 4000, "gloo"), (3, 32000, "gloo"), (2, 32000, "nccl")])
def test_decomposed_run_matches_single_gpu(tmp_path, world, n, backend):
    """backend gloo: all ranks share cuda:0 and stage exchanges through host memory (runs on a one-GPU box);
    backend nccl: one GPU per rank, device-to-device exchange over NVLink (skipped below two GPUs)."""
    import torch
    if backend == "nccl" and torch.cuda.device_count() < world:
        pytest.skip("needs one GPU per rank")
    steps = 600
    out = str(tmp_path / "dd.npz")
    launch("gpu_parity", world, out, {"DD_N": str(n), "DD_STEPS": str(steps), "DD_BACKEND": backend}, timeout=600)
    r = np.load(out)
    a, b, updates = single_gpu_reference(n, steps)
    # initial state: same positions => identical neighbour sets (by particle id) and forces to 1e-6
    # (the t0 forces come from different kernel variants - with / without auxiliary variables - hence no bit test)
    sets, rel = compare(a, r, "t0", n, bitwise=(world == 1), bitwise_force=False)
    assert all(np.array_equal(x, y) for x, y in zip(a["sets"], sets)), "neighbour sets differ from the single-GPU run"
    assert rel < 1e-6, rel
    assert np.allclose(a["thermo"], r["t0_thermo"], rtol=2e-6, atol=1e-6)
    # after the run (several list updates with migration): one rank reproduces the single-GPU trajectory bit for bit;
    # several ranks differ only by the summation order of the forces
    sets, rel = compare(b, r, "t1", n, bitwise=(world == 1), bitwise_force=(world == 1))
    assert int(r["updates"][0]) >= 2
    assert int(r["stale_ghosts"][0]) == 0                  # halo pushes over peer memory: ghosts are the owners' current positions
    if world > 1:
        assert int(r["migrated"][0]) > 0
        L = (n / 0.8442) ** (1 / 3)
        d = b["position"] - r["t1_position"]
        d -= np.rint(d / L) * L
        assert np.abs(d).max() < 1e-3                      # chaotic divergence from ulp-level differences stays tiny
        same = sum(np.array_equal(x, y) for x, y in zip(b["sets"], sets))
        assert same > 0.98 * n                             # sets may differ for pairs at the rounding edge of r_c + skin
        assert rel < 1e-2
        # Newton's third law across the slab faces: both ranks must have used the same positions for every cross-boundary
        # pair in every step, or the total momentum would drift (a ghost one step stale shifts v_cm by ~1e-6 per step)
        assert np.abs(r["t1_vcm"] - r["t0_vcm"]).max() < 2e-7, r["t1_vcm"] - r["t0_vcm"]
        assert abs(r["t1_thermo"][0] - r["t0_thermo"][0]) < 2e-5 * abs(r["t0_thermo"][0])      # total energy conserved
    else:
        assert all(np.array_equal(x, y) for x, y in zip(b["sets"], sets))
    assert np.allclose(b["thermo"][:2], r["t1_thermo"][:2], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_decomposed_config4_size_one_rank(cuda):
    """BASELINE config 4 at its per-GPU size (8 M particles) through the decomposed step loop on one rank: Andersen
    melt, then NVE.  Size-independent properties: ids survive migration-free re-orderings, energy is conserved, the total
    momentum stays zero, lists are rebuilt on the displacement criterion, the force kernel walks private shell rows."""
    from halmd_b200 import simulation
    n = 8000000
    dd = simulation.decomposed_lj_fluid(n, 0, 1, seed=1, integrator="andersen", timestep=0.005, rate=2.0)
    dd.run(60)
    assert 0.5 < dd.thermodynamics()["temperature"] < 1.5          # (the lattice is still melting)
    dd.set_integrator("verlet", timestep=0.001)
    th0 = dd.thermodynamics()
    v0 = dd.download()["velocity"].mean(axis=0)
    updates0 = dd.updates
    dd.run(120)
    th1 = dd.thermodynamics()
    assert abs(th1["internal_energy"] - th0["internal_energy"]) < 1e-5 * abs(th0["internal_energy"])
    assert dd.updates - updates0 >= 2 and dd.shells is not None
    st = dd.local_state()
    assert np.array_equal(np.sort(st["id"]), np.arange(n))
    assert np.all(np.abs(st["velocity"].mean(axis=0) - v0) < 1e-7)       # unit masses: the mean velocity is the momentum
    L = dd.L
    assert np.all(np.abs(st["position"]) <= 0.5 * L * (1 + 1e-6))
