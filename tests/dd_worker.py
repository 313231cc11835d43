"""Worker process of the multi-rank decomposition tests (launched by tests/test_decomposition*.py, one per rank).

    python tests/dd_worker.py <mode> <outfile>        with RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the env

mode cpu_protocol : host logic only (gloo, CPU tensors): migration + ghost protocol on numpy particle sets
mode gpu_parity   : DecomposedSystem on cuda:0 for every rank (gloo stages through host memory), results gathered on
                    rank 0 and written to <outfile> (npz) for comparison with the single-GPU run
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_protocol(rank, world, out):
    import torch
    import torch.distributed as dist
    from halmd_b200.decomposition import Exchange, SlabDecomposition

    L = np.array([24.0, 11.0, 13.0])
    halo = 3.0
    n = 6000
    rng = np.random.default_rng(5)                       # same stream on all ranks: a global configuration
    pos = ((rng.random((n, 3)) - 0.5) * L).astype(np.float32)
    ids = np.arange(n)
    dec = SlabDecomposition(L, world, rank, halo)
    ex = Exchange(rank, world)
    mine = dec.slab_of(pos[:, 0]) == rank
    p, i = pos[mine].copy(), ids[mine].copy()
    checks = {}
    for it in range(4):
        # random displacement below one slab width, wrapped
        d = np.random.default_rng(100 + it).normal(scale=1.2, size=(n, 3)).astype(np.float32)
        p = p + d[i]
        p = (p + L.astype(np.float32) / 2) % L.astype(np.float32) - L.astype(np.float32) / 2
        # migration, as DecomposedSystem._migrate: stable partition by key, exchange counts, exchange records
        key = dec.owner_keys(p[:, 0])
        assert not np.any(key == 3)
        order = np.argsort(key, kind="stable")
        p, i, key = p[order], i[order], key[order]
        n_stay, n_l, n_r = (int((key == k).sum()) for k in (0, 1, 2))
        fr, fl = ex.exchange_counts(n_l, n_r)
        send_l = torch.from_numpy(np.ascontiguousarray(p[n_stay:n_stay + n_l]))
        send_r = torch.from_numpy(np.ascontiguousarray(p[n_stay + n_l:]))
        recv_r, recv_l = torch.zeros(fr, 3), torch.zeros(fl, 3)
        ex.sendrecv(send_l, send_r, recv_r, recv_l)
        isend_l = torch.from_numpy(np.ascontiguousarray(i[n_stay:n_stay + n_l]))
        isend_r = torch.from_numpy(np.ascontiguousarray(i[n_stay + n_l:]))
        irecv_r, irecv_l = torch.zeros(fr, dtype=torch.int64), torch.zeros(fl, dtype=torch.int64)
        ex.sendrecv(isend_l, isend_r, irecv_r, irecv_l)
        p = np.concatenate([p[:n_stay], recv_r.numpy(), recv_l.numpy()])
        i = np.concatenate([i[:n_stay], irecv_r.numpy(), irecv_l.numpy()])
        assert np.all(dec.slab_of(p[:, 0]) == rank), "a particle arrived at the wrong rank"
        # ghosts
        gp, gi = [], []
        cnt = {"left": 0, "right": 0}
        buf = {"left": torch.zeros(0, 3), "right": torch.zeros(0, 3)}
        ibuf = {"left": torch.zeros(0, dtype=torch.int64), "right": torch.zeros(0, dtype=torch.int64)}
        for side, dest in dec.ghost_sides():
            m = dec.ghost_mask(p[:, 0], side)
            cnt[dest] = int(m.sum())
            buf[dest] = torch.from_numpy(np.ascontiguousarray(p[m]))
            ibuf[dest] = torch.from_numpy(np.ascontiguousarray(i[m]))
        fr, fl = ex.exchange_counts(cnt["left"], cnt["right"])
        g_r, g_l = torch.zeros(fr, 3), torch.zeros(fl, 3)
        ex.sendrecv(buf["left"], buf["right"], g_r, g_l)
        gi_r, gi_l = torch.zeros(fr, dtype=torch.int64), torch.zeros(fl, dtype=torch.int64)
        ex.sendrecv(ibuf["left"], ibuf["right"], gi_r, gi_l)
        ghost_p = np.concatenate([g_r.numpy(), g_l.numpy()])
        ghost_i = np.concatenate([gi_r.numpy(), gi_l.numpy()])
        # global view for the check (all_gather_object is fine for a few thousand particles)
        allp, alli = [None] * world, [None] * world
        dist.all_gather_object(allp, p)
        dist.all_gather_object(alli, i)
        gp_all, gi_all = np.concatenate(allp), np.concatenate(alli)
        assert len(gi_all) == n and len(np.unique(gi_all)) == n, "particles lost or duplicated in migration"
        assert len(np.unique(ghost_i)) == len(ghost_i), "a ghost arrived twice"
        assert not np.any(np.isin(ghost_i, i)), "a ghost duplicates an owned particle"
        # completeness: every foreign particle within `halo` (minimum image) of an owned one must be a ghost here
        foreign = ~np.isin(gi_all, i)
        fp, fi = gp_all[foreign], gi_all[foreign]
        Lf = L.astype(np.float32)
        need = np.zeros(len(fi), dtype=bool)
        for a in range(0, len(p), 512):
            dvec = p[a:a + 512, None, :] - fp[None, :, :]
            dvec -= np.rint(dvec / Lf) * Lf
            need |= ((dvec ** 2).sum(-1) <= halo ** 2).any(axis=0)
        missing = np.setdiff1d(fi[need], ghost_i)
        assert len(missing) == 0, f"rank {rank}: {len(missing)} ghosts missing"
        # ghost positions are the owners' current positions
        lookup = {int(k): v for k, v in zip(gi_all, gp_all)}
        assert all(np.array_equal(lookup[int(k)], v) for k, v in zip(ghost_i, ghost_p))
        checks[f"iter{it}"] = (len(p), len(ghost_i), n_l + n_r)
    # reductions
    t = torch.tensor([float(rank + 1)])
    ex.allreduce(t, "sum")
    assert t.item() == world * (world + 1) / 2
    t = torch.tensor([rank * 3], dtype=torch.int32)
    ex.allreduce(t, "max")
    assert t.item() == 3 * (world - 1)
    if rank == 0:
        np.savez(out, **{k: np.array(v) for k, v in checks.items()})


def gpu_parity(rank, world, out):
    import torch
    from halmd_b200 import simulation
    from halmd_b200.decomposition import DecomposedSystem, SlabDecomposition
    import torch.distributed as dist

    torch.cuda.set_device(rank if os.environ.get("DD_BACKEND") == "nccl" else 0)
    n = int(os.environ.get("DD_N", "4000"))
    steps = int(os.environ.get("DD_STEPS", "100"))
    rho = 0.8442
    L = np.array([(n / rho) ** (1 / 3)] * 3)
    rng = np.random.default_rng(21)
    pos = simulation.fcc_positions(n, L)[rng.permutation(n)].astype(np.float32).astype(np.float64)
    vel = simulation.boltzmann_velocities(n, 1.2, rng)
    ids = np.arange(n)
    dec = SlabDecomposition(L, world, rank, 3.0)
    mine = dec.slab_of(pos[:, 0].astype(np.float32)) == rank
    s = DecomposedSystem(L, pos[mine], vel[mine], ids[mine], n_global=n, rank=rank, nranks=world)
    res = {}

    def snapshot(tag):
        st = s.local_state()
        sets = s.neighbour_sets()
        th = s.thermodynamics()
        gathered = [None] * world
        dist.all_gather_object(gathered, (st, sets))
        if rank == 0:
            idx = np.concatenate([g[0]["id"] for g in gathered])
            order = np.argsort(idx)
            assert np.array_equal(idx[order], np.arange(n)), "ownership is not a partition"
            for key in ("position", "velocity", "force"):
                res[f"{tag}_{key}"] = np.concatenate([g[0][key] for g in gathered])[order]
            allsets = {}
            for g in gathered:
                allsets.update(g[1])
            res[f"{tag}_nl_count"] = np.array([len(allsets[i]) for i in range(n)])
            res[f"{tag}_nl_flat"] = np.concatenate([allsets[i] for i in range(n)])
            res[f"{tag}_thermo"] = np.array([th["internal_energy"], th["potential_energy"], th["temperature"], th["pressure"]])
            res[f"{tag}_vcm"] = np.array(th["center_of_mass_velocity"])

    def ghost_check():
        """Ghost slots must hold, bit for bit, the owners' current positions (the halo pushes of the last step)."""
        s._finish_step()
        torch.cuda.synchronize()
        ntot = s.n_own + s.n_ghost
        ids = s.id[:ntot].cpu().numpy().astype(np.int64)
        pos = s.pos[:ntot].cpu().numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, (ids[:s.n_own], pos[:s.n_own]))
        owner_pos = np.zeros((n, 4), dtype=np.float32)
        for i, p in gathered:
            owner_pos[i] = p
        bad = int((pos[s.n_own:].view(np.uint32) != owner_pos[ids[s.n_own:]].view(np.uint32)).any(axis=1).sum())
        counts = [None] * world
        dist.all_gather_object(counts, bad)
        return int(sum(counts))

    snapshot("t0")
    s.run(steps // 2)
    stale = ghost_check()
    s.run(steps - steps // 2)
    stale += ghost_check()
    snapshot("t1")
    if rank == 0:
        res["stale_ghosts"] = np.array([stale])
        res["updates"] = np.array([s.updates])
        res["migrated"] = np.array([s.stats["migrated"]])
        np.savez(out, **res)


def main():
    import torch.distributed as dist
    mode, out = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    backend = os.environ.get("DD_BACKEND", "gloo")
    if backend == "nccl":
        import torch
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        {"cpu_protocol": cpu_protocol, "gpu_parity": gpu_parity}[mode](rank, world, out)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
