"""Slab domain decomposition of the MD step loop across the GPUs of one node (one process per GPU).

The reference has no multi-GPU mode (doc/usage/multi-gpu.rst.in runs independent replicas); this module is new
(SURVEY §8e).  Rank r of P owns the particles whose wrapped x coordinate lies in
[-L_x/2 + r*W, -L_x/2 + (r+1)*W), W = L_x / P, and keeps ghost copies of the particles of its two neighbours that lie
within r_c + skin of its faces.  Particle coordinates stay global, so the kernels' minimum-image convention handles
the periodic wrap, and binning uses the *global* cell grid of the single-GPU run: every owned particle therefore gets
exactly the neighbour set the single-GPU code would give it.

Per MD step (hx_dd_enqueue_steps, csrc/dd.cu — no host code and no library collective between the kernels of a step):
               fused finalize/integrate of the owned particles into the other position buffer -> the boundary
               particles' new positions stored straight into the neighbours' ghost slots (peer memory over NVLink) ->
               signal {step, local displacement maximum} to every rank's inbox -> wait for all inboxes, global maximum
               into the control block (the rebuild vote) -> guarded pair force (owned rows, ghost partners)
Per list update (when the device-side displacement test trips): migrate leavers to the neighbouring ranks, Hilbert
order, select + exchange ghosts, bin, build lists, reset displacement origin — host-orchestrated, torch.distributed.

Three layers, so that the host logic is testable without a GPU:
  SlabDecomposition  geometry and ownership arithmetic (pure numpy restatement of csrc/dd.cu, used by the tests)
  Exchange           neighbour exchange / reductions over torch.distributed (NCCL on device tensors; gloo stages
                     through host memory, which is what the CPU tests and the single-GPU two-rank test use)
  DecomposedSystem   device state + step loop through the C ABI (needs CUDA; no CPU fallback)
"""
import ctypes as C
import math

import numpy as np

__all__ = ["SlabDecomposition", "Exchange", "DecomposedSystem"]


class SlabDecomposition:
    """Geometry of the slab decomposition along x; float32 arithmetic identical to csrc/dd.cu."""

    def __init__(self, box_length, nranks, rank, halo):
        self.L = np.asarray(box_length, dtype=np.float64)
        self.nranks, self.rank = int(nranks), int(rank)
        if not (0 <= self.rank < self.nranks):
            raise ValueError("rank out of range")
        self.x0 = np.float32(-self.L[0] / 2)
        self.width = np.float32(self.L[0] / self.nranks)
        self.inv_width = np.float32(1.0) / self.width
        # a little slack absorbs the rounding differences between slab_of() and the face coordinates
        self.halo = np.float32(float(halo) * (1 + 1e-5) + 1e-4)
        self.lo = np.float32(self.x0 + np.float32(self.rank) * self.width)
        self.hi = np.float32(self.lo + self.width)
        self.left = (self.rank - 1) % self.nranks
        self.right = (self.rank + 1) % self.nranks
        if self.nranks > 1:
            if self.nranks == 2:
                if self.L[0] / 2 < float(self.halo):
                    raise ValueError("box too small for two slabs: L_x / 2 < r_c + skin")
            elif float(self.width) < float(self.halo):
                raise ValueError("slab width below r_c + skin: use fewer ranks")

    def cell_layers(self, cell_length, ncell_x):
        """Layers of the global cell grid in x that can hold this rank's particles (owned and ghosts): first layer and
        count, periodic.  The cell of x is uint(x / l + n) % n (csrc/binning.cu); one layer of margin on either side
        covers the rounding of that quotient and of the slab faces."""
        if self.nranks == 1:
            return 0, int(ncell_x)
        a = int(math.floor((float(self.lo) - float(self.halo)) / cell_length + ncell_x)) - 1
        b = int(math.floor((float(self.hi) + float(self.halo)) / cell_length + ncell_x)) + 1
        count = b - a + 1
        if count >= ncell_x:
            return 0, int(ncell_x)
        return a % int(ncell_x), count

    def slab_of(self, x):
        """Owner slab of wrapped coordinates x (float32), as k_dd_owner_keys computes it."""
        x = np.asarray(x, dtype=np.float32)
        s = np.floor((x - self.x0) * self.inv_width).astype(np.int64)
        return np.mod(s, self.nranks)

    def owner_keys(self, x):
        """0 stays, 1 to the left neighbour, 2 to the right neighbour (1 for both when P = 2), 3 lost."""
        s = self.slab_of(x)
        key = np.full(s.shape, 3, dtype=np.uint32)
        key[s == self.right] = 2
        key[s == self.left] = 1
        key[s == self.rank] = 0
        if self.nranks == 2:
            key[key == 2] = 1
        return key

    def ghost_mask(self, x, side):
        """Particles within the halo of the low (0), high (1) or either (2) face."""
        x = np.asarray(x, dtype=np.float32)
        # distances to the faces modulo the box length, as k_dd_ghost_keys folds them: a coordinate of +L/2 belongs to
        # the slab that starts at -L/2
        Lx = np.float32(self.L[0])
        d_lo, d_hi = x - self.lo, self.hi - x
        half = np.float32(0.5) * Lx          # a period away — not a rounding error of the face coordinates
        up = d_lo > half
        down = ~up & (d_hi > half)
        d_lo = np.where(up, d_lo - Lx, np.where(down, d_lo + Lx, d_lo)).astype(np.float32)
        d_hi = np.where(up, d_hi + Lx, np.where(down, d_hi - Lx, d_hi)).astype(np.float32)
        near_lo, near_hi = d_lo < self.halo, d_hi < self.halo
        return near_lo if side == 0 else (near_hi if side == 1 else (near_lo | near_hi))

    def ghost_sides(self):
        """Ghost messages of a list update as (side, destination): one combined list for two ranks (both faces
        border the same peer, which must receive every particle once), low/high lists otherwise."""
        if self.nranks == 1:
            return []
        if self.nranks == 2:
            return [(2, "left")]
        return [(0, "left"), (1, "right")]


class Exchange:
    """Neighbour exchange along the slab direction and small reductions, on torch.distributed.

    A "leftward" message travels to rank-1 and is received from rank+1; a "rightward" one the other way round.  With
    two ranks both directions connect the same pair, so operations are always issued leftward first (NCCL matches
    point-to-point operations per peer in issue order; gloo additionally gets distinct tags)."""

    def __init__(self, rank=0, nranks=1, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.nranks, self.group = rank, nranks, group
        self.left, self.right = (rank - 1) % nranks, (rank + 1) % nranks
        self.backend = dist.get_backend(group) if (nranks > 1 and dist.is_initialized()) else None
        self.stage = self.backend == "gloo"        # gloo moves host memory: device tensors are staged
        self.bytes_sent = 0

    def _wire(self, t):
        return t.cpu() if (self.stage and t.is_cuda) else t

    def sendrecv(self, leftward=None, rightward=None, from_right=None, from_left=None):
        """Send `leftward` to rank-1 and `rightward` to rank+1; fill `from_right` with the leftward message of rank+1
        and `from_left` with the rightward message of rank-1.  Tensors may be None or empty (skipped on both ends:
        the caller exchanges the sizes first)."""
        if self.nranks == 1:
            return
        dist, ops, post = self.dist, [], []

        def send(t, peer, tag):
            if t is not None and t.numel() > 0:
                w = self._wire(t.contiguous())
                self.bytes_sent += w.numel() * w.element_size()
                ops.append(dist.P2POp(dist.isend, w, peer, self.group, tag))

        def recv(t, peer, tag):
            if t is not None and t.numel() > 0:
                if self.stage and t.is_cuda:
                    w = self.torch.empty(t.shape, dtype=t.dtype, device="cpu")
                    post.append((t, w))
                else:
                    assert t.is_contiguous()
                    w = t
                ops.append(dist.P2POp(dist.irecv, w, peer, self.group, tag))

        send(leftward, self.left, 0)
        recv(from_right, self.right, 0)
        send(rightward, self.right, 1)
        recv(from_left, self.left, 1)
        if ops:
            for work in dist.batch_isend_irecv(ops):
                work.wait()
        for t, w in post:
            t.copy_(w)

    def sendrecv_many(self, items):
        """Several (leftward, rightward, from_right, from_left) exchanges issued as ONE batch of point-to-point operations."""
        if self.nranks == 1:
            return
        dist, ops, post = self.dist, [], []
        for k, (leftward, rightward, from_right, from_left) in enumerate(items):
            for t, peer, tag, is_send in ((leftward, self.left, 2 * k, True), (from_right, self.right, 2 * k, False),
                                          (rightward, self.right, 2 * k + 1, True), (from_left, self.left, 2 * k + 1, False)):
                if t is None or t.numel() == 0:
                    continue
                if is_send:
                    w = self._wire(t.contiguous())
                    self.bytes_sent += w.numel() * w.element_size()
                    ops.append(dist.P2POp(dist.isend, w, peer, self.group, tag))
                else:
                    if self.stage and t.is_cuda:
                        w = self.torch.empty(t.shape, dtype=t.dtype, device="cpu")
                        post.append((t, w))
                    else:
                        assert t.is_contiguous()
                        w = t
                    ops.append(dist.P2POp(dist.irecv, w, peer, self.group, tag))
        if ops:
            for work in dist.batch_isend_irecv(ops):
                work.wait()
        for t, w in post:
            t.copy_(w)

    def allgather_ints(self, values):
        """values (a few Python ints) of every rank: list of lists, one collective and one host read."""
        if self.nranks == 1:
            return [list(values)]
        torch = self.torch
        dev = "cpu" if self.stage else "cuda"
        mine = torch.tensor(list(values), dtype=torch.int64, device=dev)
        out = torch.empty(self.nranks * len(values), dtype=torch.int64, device=dev)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        return out.cpu().numpy().reshape(self.nranks, len(values)).tolist()

    def exchange_counts(self, n_leftward, n_rightward):
        """Sizes of the messages about to arrive: (from_right, from_left)."""
        if self.nranks == 1:
            return 0, 0
        torch = self.torch
        dev = "cpu" if self.stage else "cuda"
        a = torch.tensor([n_leftward], dtype=torch.int64, device=dev)
        b = torch.tensor([n_rightward], dtype=torch.int64, device=dev)
        ra, rb = torch.zeros_like(a), torch.zeros_like(b)
        self.sendrecv(a, b, ra, rb)
        return int(ra.item()), int(rb.item())

    def allreduce(self, t, op="sum"):
        if self.nranks == 1:
            return t
        dist = self.dist
        o = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX}[op]
        if self.stage and t.is_cuda:
            w = t.cpu()
            dist.all_reduce(w, op=o, group=self.group)
            t.copy_(w)
        else:
            dist.all_reduce(t, op=o, group=self.group)
        return t

    def barrier(self):
        if self.nranks > 1:
            self.dist.barrier(group=self.group)

    def allgather_object(self, obj):
        """[obj of rank 0, ..., obj of rank P-1] (small host objects: IPC handles, buffer indices)."""
        if self.nranks == 1:
            return [obj]
        out = [None] * self.nranks
        self.dist.all_gather_object(out, obj, group=self.group)
        return out


def _round_up(n, m):
    return ((int(n) + m - 1) // m) * m


class _DeviceArray:
    """A device allocation of the C library exposed through the CUDA array interface (so torch can view it)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2}
        self.owner = owner


def _device_view(torch, ptr, shape, typestr, owner):
    return torch.as_tensor(_DeviceArray(ptr, shape, typestr, owner), device="cuda")


class DecomposedSystem:
    """One rank's share of a slab-decomposed LJ system: device state, neighbour structures and the step loop.

    Arrays have `cap` slots: owned particles in [0, n_own), ghosts in [n_own, n_own + n_ghost).  All kernels are the
    single-GPU ones, called through the C ABI with explicit counts (hx_*2 / hx_*_range / hx_dd_*)."""

    def __init__(self, box_length, positions, velocities, ids, species=None, nspecies=1, epsilon=(1.0,), sigma=(1.0,),
                 cutoff=(2.5,), truncation="smooth_r4", h=0.005, skin=0.5, timestep=0.001, integrator="verlet",
                 temperature=1.2, rate=2.0, seed=1, n_global=None, rank=0, nranks=1, group=None, occupancy=None,
                 lanes=4, capacity=None, sorting=True):
        import torch
        from . import _lib as hx
        if not torch.cuda.is_available():
            raise hx.HxError("DecomposedSystem needs a CUDA device (halmd_b200 has no CPU fallback)")
        self.torch, self.hx, self.lib = torch, hx, hx.lib
        self.rank, self.nranks = rank, nranks
        self.L = np.asarray(box_length, dtype=np.float64)
        self.Lf = hx.f3(self.L)
        self.stream = torch.cuda.current_stream()
        self.st = C.c_void_p(self.stream.cuda_stream)
        self.dt = float(timestep)
        self.skin = float(skin)
        self.skin_half = float(np.float32(np.float32(skin) / 2))
        self.kind = integrator
        self.step = 0
        self.updates = 0
        self.sorting = sorting
        self.stats = {"chunks": 0, "syncs": 0, "migrated": 0}
        self.lanes_flag = hx.FORCE_LANES(lanes)

        nt = int(nspecies)
        trunc_map = {"sharp": hx.TRUNC_SHARP, "shifted": hx.TRUNC_SHIFTED, "smooth_r4": hx.TRUNC_SMOOTH_R4,
                     "force_shifted": hx.TRUNC_FORCE_SHIFTED}
        if truncation not in trunc_map:
            raise hx.HxError(f"unknown truncation '{truncation}'")
        trunc = trunc_map[truncation]
        self.pot = hx.LjParams()
        hx.check(self.lib.hx_potential_lj_init(C.byref(self.pot), nt, nt, hx.farr(epsilon), hx.farr(sigma), trunc,
                                               hx.farr(cutoff), float(h)), "potential")
        r_cut = np.array([self.pot.r_cut[i] for i in range(nt * nt)], dtype=np.float32)
        self.nt = nt
        self.r_cut_max = float(r_cut.max())
        self.rr_cut_skin = hx.farr(((r_cut + np.float32(skin)).astype(np.float64) ** 2).astype(np.float32))
        self.dec = SlabDecomposition(self.L, nranks, rank, self.r_cut_max + skin)
        self.ex = Exchange(rank, nranks, group)

        n_own = int(positions.shape[0])
        self.n_global = int(n_global if n_global is not None else n_own)
        # global cell grid and list size: identical to the single-GPU plan (binning.cpp:81-101, from_binning.cpp:96-119)
        ncell, cl, cs = (C.c_uint * 3)(), (C.c_float * 3)(), C.c_uint()
        occ_b = float(occupancy) if occupancy else 0.5
        occ_n = float(occupancy) if occupancy else 0.4
        hx.check(self.lib.hx_binning_plan(hx.d3(self.L), self.n_global, self.r_cut_max, float(skin), occ_b, ncell, cl, C.byref(cs)))
        self.ncell, self.cell_length, self.cell_size = ncell, cl, cs.value
        if min(ncell) < 3:
            raise hx.HxError("neighbour: number of cells per dimension must be at least 3")
        self.ncells = int(np.prod(list(ncell)))
        self.cell_x0, self.cell_nx = self.dec.cell_layers(float(cl[0]), int(ncell[0]))
        size = C.c_uint()
        hx.check(self.lib.hx_nlist_plan(self.r_cut_max, float(skin), self.n_global, float(np.prod(self.L)), occ_n, self.cell_size, C.byref(size)))
        self.occ_n = occ_n
        self.size = size.value
        self.stride = ((self.size + 32) // 32) * 32
        self.depth = self.lib.hx_hilbert_depth(hx.d3(self.L))

        # capacity: owned share with head room + two ghost layers
        share = self.n_global / nranks
        ghost_frac = 0.0 if nranks == 1 else min(1.0, 2.0 * float(self.dec.halo) / float(self.dec.width))
        if capacity is None:
            own_cap = _round_up(max(n_own, share) * 1.08 + 4096, 1024)
            capacity = _round_up(own_cap + share * ghost_frac * 1.15 + 4096, 1024)
        else:
            own_cap = _round_up(min(capacity, max(n_own, share) * 1.08 + 4096), 1024)
        self.cap, self.own_cap = int(capacity), int(own_cap)

        dev = torch.device("cuda", torch.cuda.current_device())
        f32, i32 = torch.float32, torch.int32
        z = lambda *shape, dtype=f32: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        cap = self.cap
        # positions: two buffers allocated by the library (plain cudaMalloc: exportable to the other ranks), used
        # alternately by the step loop; `cur` is the one holding the current positions
        self._pos_ptr, self._pos_t, self.cur = [], [], 0
        for _ in range(2):
            ptr = C.c_void_p()
            hx.check(self.lib.hx_dd_alloc(2 * cap * 16, C.byref(ptr)), "hx_dd_alloc")
            self._pos_ptr.append(ptr.value)
            self._pos_t.append(_device_view(torch, ptr.value, (2 * cap, 4), "<f4", self))
        inbox = C.c_void_p()
        hx.check(self.lib.hx_dd_alloc(2 * hx.DD_MAX_RANKS * 8, C.byref(inbox)), "hx_dd_alloc")
        self._inbox_ptr = inbox.value
        self._map_peers()
        self.vel, self.vel_b = z(2 * cap, 4), z(2 * cap, 4)
        self.image, self.image_b = z(cap, 4), z(cap, 4)
        self.force = z(cap, 4)
        self.id, self.id_b = z(cap, dtype=i32), z(cap, dtype=i32)
        self.en_pot, self.stress = z(cap), z(6 * cap)
        self.r0 = z(self.own_cap, 4)
        self.keys, self.index = z(cap, dtype=i32), z(cap, dtype=i32)
        self.sort_bytes = self.lib.hx_sort_scratch_bytes(cap)
        self.sort_scratch = torch.empty(self.sort_bytes, dtype=torch.uint8, device=dev)
        self.counts = z(4, dtype=i32)
        self.ctl = z(4, dtype=i32)
        self.flags = z(4, dtype=i32)
        # placeholders everywhere, once: list updates rewrite this slab's layers only (hx_binning_update_slab), and no
        # particle of this rank ever maps outside them (checked on the device)
        self.g_cell = torch.full((self.ncells * self.cell_size,), -1, dtype=i32, device=dev)
        self.cell_count = z(self.ncells, dtype=i32)
        self.nl = torch.empty(self.own_cap * self.stride, dtype=i32, device=dev)
        self.nl_count = z(self.own_cap, dtype=i32)
        self.thermo = torch.zeros(hx.THERMO_DOUBLES, dtype=torch.float64, device=dev)
        self.send_idx = {"left": z(0, dtype=i32), "right": z(0, dtype=i32)}
        self.send_buf = {"left": z(0, 4), "right": z(0, 4)}
        self.n_ghost_from = {"left": 0, "right": 0}

        # upload the owned particles (double-single split as dsdeq, dsfun.hpp:122-126)
        self.n_own, self.n_ghost = n_own, 0
        if n_own > self.own_cap:
            raise hx.HxError("initial share exceeds the capacity")
        pos = np.asarray(positions, dtype=np.float64)
        vel = np.asarray(velocities, dtype=np.float64)
        sp = np.zeros(n_own, dtype=np.uint32) if species is None else np.asarray(species, dtype=np.uint32)
        ph = np.zeros((n_own, 4), dtype=np.float32)
        pl = np.zeros((n_own, 4), dtype=np.float32)
        ph[:, :3] = pos
        pl[:, :3] = pos - ph[:, :3].astype(np.float64)
        ph[:, 3] = sp.view(np.float32)
        vh = np.zeros((n_own, 4), dtype=np.float32)
        vl = np.zeros((n_own, 4), dtype=np.float32)
        vh[:, :3] = vel
        vl[:, :3] = vel - vh[:, :3].astype(np.float64)
        vh[:, 3] = 1.0
        self.pos[:n_own] = torch.from_numpy(ph).to(dev)
        self.pos[cap:cap + n_own] = torch.from_numpy(pl).to(dev)
        self.vel[:n_own] = torch.from_numpy(vh).to(dev)
        self.vel[cap:cap + n_own] = torch.from_numpy(vl).to(dev)
        self.id[:n_own] = torch.from_numpy(np.asarray(ids, dtype=np.int64).astype(np.int32)).to(dev)

        if integrator == "andersen":
            self.rng = hx.Rand48()
            self.rng_state = torch.zeros(3 * 32 * 256, dtype=torch.int16, device=dev)
            hx.check(self.lib.hx_rand48_seed(C.byref(self.rng), 32, 256, int(seed) + 7919 * rank, self._p(self.rng_state), self.st))
            self.set_temperature(temperature, rate)
        self.use_shells = True
        self._staging = {}
        self.shells, self.shell_rows, self.shell_ends = None, None, None
        import os
        self._profile, self._tick_t = bool(os.environ.get("HALMD_B200_DD_PROFILE")), 0.0
        self.timeout_cycles = int(20e9)            # ~10 s of SM clock: a rank that never arrives stops the chunk
        self.lists_valid = False
        self.force_valid = False
        self.pending_finalize = False
        self.last_interval, self.since_update = 16, 0
        self.min_chunk, self.max_chunk = 2, 64

    # ---- position buffers and peer memory --------------------------------------------------------------------
    @property
    def pos(self):
        return self._pos_t[self.cur]

    @property
    def pos_b(self):
        return self._pos_t[self.cur ^ 1]

    def _map_peers(self):
        """Exchange the IPC handles of the position buffers and the inbox; map those of the other ranks."""
        hx, torch = self.hx, self.torch
        self._peer_pos = {self.rank: list(self._pos_ptr)}
        self._peer_inbox = {self.rank: self._inbox_ptr}
        if self.nranks > 1:
            mine = []
            for ptr in self._pos_ptr + [self._inbox_ptr]:
                h = C.create_string_buffer(hx.DD_IPC_HANDLE_BYTES)
                hx.check(self.lib.hx_dd_ipc_export(C.c_void_p(ptr), h), "hx_dd_ipc_export")
                mine.append(h.raw)
            everyone = self.ex.allgather_object(mine)
            needed = {self.dec.left, self.dec.right}
            for r, handles in enumerate(everyone):
                if r == self.rank:
                    continue
                opened = []
                for k, raw in enumerate(handles):
                    if k < 2 and r not in needed:      # positions of the two neighbours only; every rank's inbox
                        opened.append(None)
                        continue
                    ptr = C.c_void_p()
                    hx.check(self.lib.hx_dd_ipc_open(raw, C.byref(ptr)), "hx_dd_ipc_open")
                    opened.append(ptr.value)
                self._peer_pos[r] = opened[:2]
                self._peer_inbox[r] = opened[2]
        table = np.zeros(hx.DD_MAX_RANKS, dtype=np.uint64)
        for r, ptr in self._peer_inbox.items():
            table[r] = ptr
        self.peer_inbox_table = torch.from_numpy(table.view(np.int64)).to(torch.device("cuda", torch.cuda.current_device()))
        self.peer_flip = {"left": 0, "right": 0}
        self.peer_offset = {"left": 0, "right": 0}

    def close(self):
        """Unmap the peers' buffers and free the exported ones (collective: every rank calls it)."""
        if getattr(self, "_pos_ptr", None) is None:
            return
        self.torch.cuda.synchronize()
        self.ex.barrier()
        for r, ptrs in self._peer_pos.items():
            if r != self.rank:
                for ptr in list(ptrs) + [self._peer_inbox[r]]:
                    if ptr:
                        self.lib.hx_dd_ipc_close(C.c_void_p(ptr))
        self.ex.barrier()
        self._pos_t = None
        for ptr in self._pos_ptr + [self._inbox_ptr]:
            self.lib.hx_dd_free(C.c_void_p(ptr))
        self._pos_ptr = None

    # ---- helpers ---------------------------------------------------------------------------------------------
    @staticmethod
    def _p(t, offset=0):
        return C.c_void_p(t.data_ptr() + offset * t.element_size() * (t.shape[1] if t.dim() > 1 else 1))

    def set_timestep(self, dt):
        self.dt = float(dt)
        if self.kind == "andersen":
            self.set_temperature(self.temperature, self.rate)

    def set_temperature(self, temperature, rate):
        self.temperature, self.rate = float(temperature), float(rate)
        self.sqrt_temperature = float(np.float32(math.sqrt(temperature)))
        self.coll_prob = float(np.float32(rate * self.dt))

    def set_integrator(self, kind, timestep=None, temperature=None, rate=None):
        """Swap the integrator (as constructing a new mdsim.integrators.* on the same particle instance)."""
        self._finish_step()
        if kind == "andersen" and not hasattr(self, "rng"):
            raise self.hx.HxError("construct with integrator='andersen' to allocate the RNG state")
        self.kind = kind
        if timestep is not None:
            self.dt = float(timestep)
        if kind == "andersen":
            self.set_temperature(self.temperature if temperature is None else temperature,
                                 self.rate if rate is None else rate)

    def _sort_pairs(self, n, bits):
        hx = self.hx
        hx.check(self.lib.hx_iota_u32(self._p(self.index), n, self.st))
        hx.check(self.lib.hx_sort_pairs_stable_u32(self._p(self.keys), self._p(self.index), n, bits,
                                                   self._p(self.sort_scratch), self.sort_bytes, self.st), "sort")

    def _permute(self, n):
        """Gather the first n particles by self.index into the back buffers and swap."""
        self.hx.check(self.lib.hx_dd_permute(self._p(self.index), n, self.cap, self._p(self.pos), self._p(self.pos_b),
                                             self._p(self.vel), self._p(self.vel_b), self._p(self.image), self._p(self.image_b),
                                             self._p(self.id), self._p(self.id_b), self.st), "permute")
        self.cur ^= 1
        self.vel, self.vel_b = self.vel_b, self.vel
        self.image, self.image_b = self.image_b, self.image
        self.id, self.id_b = self.id_b, self.id

    # ---- list update: migration, order, ghosts, cells, lists -------------------------------------------------------
    def _migrate(self):
        hx, torch, dec, cap = self.hx, self.torch, self.dec, self.cap
        if self.nranks == 1:
            return
        n = self.n_own
        hx.check(self.lib.hx_dd_owner_keys(self._p(self.pos), n, float(dec.x0), float(dec.width), self.nranks, self.rank,
                                           self._p(self.keys), self._p(self.counts), self.st), "owner_keys")
        cnt = self.counts.cpu().numpy().astype(np.int64)
        self._tick("migrate: owner keys + count read")
        self.stats["syncs"] += 1
        if cnt[3] != 0:
            raise hx.HxError(f"rank {self.rank}: {cnt[3]} particles moved further than one slab between list updates")
        n_stay, n_l, n_r = int(cnt[0]), int(cnt[1]), int(cnt[2])
        if n_l + n_r > 0:
            self._sort_pairs(n, 8)          # stable: [stay | leftward | rightward], memory order kept within each
            self._permute(n)
        self._tick("migrate: sort + permute")
        from_right, from_left = self.ex.exchange_counts(n_l, n_r)
        self._tick("migrate: exchange counts")
        n_new = n_stay + from_right + from_left
        if n_new > self.own_cap:
            raise hx.HxError(f"rank {self.rank}: owned particles {n_new} exceed the capacity {self.own_cap}")
        a, b = n_stay, n_stay + n_l           # send blocks [a, b) leftward and [b, n) rightward
        c, d = n_stay, n_stay + from_right    # arrivals: [c, d) from the right, [d, n_new) from the left
        # one message per direction: the full record of a particle (position and velocity words, image, id) as 21 floats;
        # arrivals overwrite the send blocks, so the outgoing records are packed first
        pos, vel = self.pos, self.vel

        def pack(lo, hi):
            if hi == lo:
                return None
            return torch.cat([pos[lo:hi], pos[cap + lo:cap + hi], vel[lo:hi], vel[cap + lo:cap + hi], self.image[lo:hi],
                              self.id[lo:hi].view(torch.float32).unsqueeze(1)], dim=1)

        def unpack(rec, lo, hi):
            if hi == lo:
                return
            pos[lo:hi], pos[cap + lo:cap + hi] = rec[:, 0:4], rec[:, 4:8]
            vel[lo:hi], vel[cap + lo:cap + hi] = rec[:, 8:12], rec[:, 12:16]
            self.image[lo:hi] = rec[:, 16:20]
            self.id[lo:hi] = rec[:, 20].contiguous().view(torch.int32)
        out_l, out_r = pack(a, b), pack(b, n)
        in_r = torch.empty(from_right, 21, dtype=torch.float32, device=pos.device) if from_right else None
        in_l = torch.empty(from_left, 21, dtype=torch.float32, device=pos.device) if from_left else None
        self.ex.sendrecv(out_l, out_r, in_r, in_l)
        unpack(in_r, c, d)
        unpack(in_l, d, n_new)
        self.stats["migrated"] += n_l + n_r
        self.n_own = n_new
        self._tick("migrate: pack, send / receive, unpack")

    def _order(self):
        """sorts.hilbert order of the owned particles (global box, so the curve is the single-GPU one)."""
        hx, n = self.hx, self.n_own
        if not self.sorting or n == 0:
            return
        hx.check(self.lib.hx_hilbert_keys(self._p(self.pos), n, self.Lf, self.depth, self._p(self.keys), self.st), "hilbert")
        self._sort_pairs(n, 3 * self.depth)
        self._permute(n)
        self._tick("order: Hilbert keys, sort, permute")

    def _ghosts(self):
        """Select the particles each neighbour needs, exchange them and append the arrivals behind the owned ones."""
        hx, torch, dec = self.hx, self.torch, self.dec
        self.n_ghost = 0
        self.n_ghost_from = {"left": 0, "right": 0}
        n = self.n_own
        counts = {"left": 0, "right": 0}
        for dest in ("left", "right"):
            self.send_idx[dest] = torch.zeros(0, dtype=torch.int32, device=self.pos.device)
            self.send_buf[dest] = torch.zeros(0, 4, dtype=torch.float32, device=self.pos.device)
        for side, dest in dec.ghost_sides():
            hx.check(self.lib.hx_dd_ghost_keys(self._p(self.pos), n, float(dec.lo), float(dec.hi), float(dec.halo), side,
                                               float(np.float32(self.L[0])), self._p(self.keys), self._p(self.counts), self.st), "ghost_keys")
            self._sort_pairs(n, 8)
            m = int(self.counts[0].item())
            self.stats["syncs"] += 1
            self.send_idx[dest] = self.index[:m].clone()
            self.send_buf[dest] = torch.empty(m, 4, dtype=torch.float32, device=self.pos.device)
            counts[dest] = m
        self._tick("ghosts: keys, sorts, count reads")
        if self.nranks == 1:
            return
        # one collective tells every rank what it needs: the sizes of the blocks about to arrive and where its own blocks
        # land in the neighbours' buffers (slot offset, which of their two position buffers is current)
        meta = self.ex.allgather_ints([n, counts["left"], counts["right"], self.cur])
        self._tick("ghosts: allgather of sizes")
        L, R = dec.left, dec.right
        if self.nranks == 2:
            from_right, from_left = meta[R][1], 0          # the peer's single combined list arrives as its leftward message
        else:
            from_right, from_left = meta[R][1], meta[L][2]
        total = n + from_right + from_left
        if total > self.cap:
            raise hx.HxError(f"rank {self.rank}: owned + ghost particles {total} exceed the capacity {self.cap}")
        self.n_ghost_from = {"right": from_right, "left": from_left}
        self.n_ghost = from_right + from_left
        # my leftward block is the left neighbour's `from_right` block (behind its owned particles); my rightward block
        # lands behind the right neighbour's owned particles and its own `from_right` block
        self.peer_offset = {"left": meta[L][0], "right": meta[R][0] + (meta[(R + 1) % self.nranks][1] if self.nranks > 2 else 0)}
        # relative orientation of my two position buffers and each neighbour's, fixed until the next list update (every
        # executed step switches the current buffer on all ranks alike)
        self.peer_flip = {"left": meta[L][3] ^ self.cur, "right": meta[R][3] ^ self.cur}
        # ghost positions and ids (diagnostics, neighbour-set parity tests) in one batch of messages
        ids = {}
        for dest in ("left", "right"):
            m = counts[dest]
            ids[dest] = torch.empty(m, dtype=torch.int32, device=self.pos.device)
            if m:
                hx.check(self.lib.hx_gather_float4(self._p(self.pos), self._p(self.send_idx[dest]), m, self._p(self.send_buf[dest]), None, self.st))
                hx.check(self.lib.hx_gather_u32(self._p(self.id), self._p(self.send_idx[dest]), m, self._p(ids[dest]), self.st))
        fr, fl = from_right, from_left
        self.ex.sendrecv_many([(self.send_buf["left"], self.send_buf["right"], self.pos[n:n + fr], self.pos[n + fr:n + fr + fl]),
                               (ids["left"], ids["right"], self.id[n:n + fr], self.id[n + fr:total])])
        self._tick("ghosts: gather, send / receive")

    def _halo_exchange(self, guard=True):
        """Current positions (high words, species tag) of the fixed ghost sets: pack, exchange, land in the ghost slots."""
        if self.nranks == 1:
            return
        hx, n = self.hx, self.n_own
        ctl = self._p(self.ctl) if guard else None
        for dest in ("left", "right"):
            m = self.send_idx[dest].shape[0]
            if m:
                hx.check(self.lib.hx_gather_float4(self._p(self.pos), self._p(self.send_idx[dest]), m, self._p(self.send_buf[dest]), ctl, self.st))
        fr, fl = self.n_ghost_from["right"], self.n_ghost_from["left"]
        self.ex.sendrecv(self.send_buf["left"], self.send_buf["right"], self.pos[n:n + fr], self.pos[n + fr:n + fr + fl])

    def _build(self):
        hx, torch = self.hx, self.torch
        ntot = self.n_own + self.n_ghost
        while True:
            self.flags.zero_()
            # placeholders, cell sorts and list blocks for this slab's layers of the global grid only
            hx.check(self.lib.hx_binning_update_slab(self._p(self.pos), ntot, self.cell_length, self.ncell, self.cell_size,
                                                     self.cell_x0, self.cell_nx, self._p(self.g_cell), self._p(self.cell_count), None,
                                                     self._p(self.flags), self.st), "binning")
            hx.check(self.lib.hx_nlist_build_slab(self._p(self.g_cell), self.ncell, self.cell_size, self.cell_x0, self.cell_nx,
                                                  self._p(self.pos), ntot, self.n_own, self.rr_cut_skin, self.nt, self.nt, self.Lf,
                                                  self._p(self.nl), self.size, self.stride, hx.LAYOUT_ROW, 0, self._p(self.nl_count),
                                                  self._p(self.flags, 1), self.st), "nlist")
            hx.check(self.lib.hx_count_max(self._p(self.nl_count), self.n_own, self._p(self.flags, 2), self.st), "count_max")
            # every rank must take the same decision: the cell grid and list size are global properties (and the longest
            # list, which sizes the private shell rows, may as well be); one reduction, one host read
            self._tick("build: binning, lists, longest list")
            self.ex.allreduce(self.flags, "max")
            f0, f1, self.longest_list = (int(v) for v in self.flags.cpu().numpy()[:3])
            self._tick("build: allreduce of flags + read")
            self.stats["syncs"] += 1
            if f0 & 2:
                # a particle outside the layers computed for this slab (never expected): report it and bin the whole grid
                # from now on — the layers are an optimisation, the global grid is always correct
                if self.cell_nx >= int(self.ncell[0]):
                    raise hx.HxError(f"rank {self.rank}: cell index outside the global grid")
                self._report_outside_layers(ntot)
                self.cell_x0, self.cell_nx = 0, int(self.ncell[0])
                self.stats["layer_fallbacks"] = self.stats.get("layer_fallbacks", 0) + 1
                continue
            if f0:
                # binning.cpp:216-219 — overcrowded cells: double the cell size
                self.cell_size *= 2
                self.g_cell = torch.full((self.ncells * self.cell_size,), -1, dtype=torch.int32, device=self.pos.device)
                continue
            if f1:
                # from_binning.cpp:277-280 — overcrowded lists: halve the occupancy
                self.occ_n /= 2
                size = C.c_uint()
                hx.check(self.lib.hx_nlist_plan(self.r_cut_max, self.skin, self.n_global, float(np.prod(self.L)), self.occ_n,
                                                self.cell_size, C.byref(size)))
                self.size = size.value
                self.stride = ((self.size + 32) // 32) * 32
                self.nl = torch.empty(self.own_cap * self.stride, dtype=torch.int32, device=self.pos.device)
                continue
            break

    def _tick(self, label):
        """Development aid (HALMD_B200_DD_PROFILE=1): device-synchronised wall time of the phases of a list update,
        accumulated over the NVE updates in stats["profile_ms"]."""
        if not self._profile:
            return
        import time
        self.torch.cuda.synchronize()
        now = time.perf_counter()
        if label is not None and self.kind == "verlet":
            prof = self.stats.setdefault("profile_ms", {})
            prof[label] = prof.get(label, 0.0) + 1e3 * (now - self._tick_t)
        self._tick_t = now

    def _report_outside_layers(self, ntot):
        import sys
        x = self.pos[:ntot, 0].float().cpu().numpy()
        ell, nx = float(self.cell_length[0]), int(self.ncell[0])
        c = np.floor(x.astype(np.float64) / ell + nx).astype(np.int64) % nx
        bad = np.nonzero((c - self.cell_x0) % nx >= self.cell_nx)[0]
        if len(bad):
            print(f"halmd_b200: rank {self.rank}: {len(bad)} particle(s) outside the cell layers [{self.cell_x0}, +{self.cell_nx}) of the "
                  f"slab [{float(self.dec.lo)}, {float(self.dec.hi)}): index {bad[:4].tolist()} (owned below {self.n_own}), "
                  f"x = {x[bad[:4]].tolist()}; binning the whole grid from now on", file=sys.stderr, flush=True)

    def update_lists(self):
        """neighbour update: on_prepend_update (migration + sort) -> binning -> lists -> displacement zero."""
        hx = self.hx
        import time
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        self._tick(None)
        self._migrate()
        t1 = time.perf_counter()
        self._order()
        self._ghosts()
        self.torch.cuda.synchronize()
        t2 = time.perf_counter()
        self._build()
        hx.check(self.lib.hx_displacement_zero(self._p(self.pos), self._p(self.r0), self.n_own, self.st))
        hx.check(self.lib.hx_ctl_reset(self._p(self.ctl), self.st))
        self._build_shells()
        self._tick("build: shell rows, displacement origin")
        if self._profile and self.kind == "verlet":
            self.stats["profile_updates"] = self.stats.get("profile_updates", 0) + 1
        self.torch.cuda.synchronize()
        t3 = time.perf_counter()
        self.stats["update_ms"] = self.stats.get("update_ms", 0.0) + 1e3 * (t3 - t0)
        self.stats["migrate_ms"] = self.stats.get("migrate_ms", 0.0) + 1e3 * (t1 - t0)
        self.stats["order_ghosts_ms"] = self.stats.get("order_ghosts_ms", 0.0) + 1e3 * (t2 - t1)
        self.stats["build_ms"] = self.stats.get("build_ms", 0.0) + 1e3 * (t3 - t2)
        self.lists_valid = True
        self.updates += 1

    def _build_shells(self):
        """Private shell-ordered rows of the owned particles for the pair-force kernel (hx_shell_rows_build)."""
        hx, torch = self.hx, self.torch
        self.shells = None
        if not self.use_shells or self.n_own == 0:
            return
        pitch = (self.longest_list + 32) // 32 * 32
        if pitch > 512:
            return
        words = int(self.lib.hx_shell_rows_words(self.own_cap, pitch))
        if self.shell_rows is None or self.shell_rows.numel() < words:
            self.shell_rows = torch.empty(words, dtype=torch.int32, device=self.vel.device)
        if self.shell_ends is None:
            self.shell_ends = torch.empty(self.own_cap * hx.NSHELL, dtype=torch.int16, device=self.vel.device)
        sh = hx.ShellRows()
        hx.check(self.lib.hx_shell_rows_plan(self.skin, C.byref(sh)))
        sh.rows_dev, sh.ends_dev, sh.pitch = self.shell_rows.data_ptr(), self.shell_ends.data_ptr(), pitch
        hx.check(self.lib.hx_shell_rows_build(C.byref(self.pot), self._p(self.nl), self.stride, self._p(self.nl_count), self._p(self.pos),
                                              self._p(self.pos), self.n_own, self.Lf, C.byref(sh), self.st), "shell_rows")
        self.shells = sh

    # ---- forces, integration ------------------------------------------------------------------------------------------
    def _force(self, aux=False, guarded=False):
        hx = self.hx
        if not aux and self.shells is not None:
            hx.check(self.lib.hx_pair_force_shells(C.byref(self.pot), self._p(self.pos), self._p(self.pos), C.byref(self.shells),
                                                   self._p(self.r0), self.n_own, self.Lf, 0, 1, self._p(self.force),
                                                   self._p(self.ctl) if guarded else None, self.skin_half, self.st), "pair_force_shells")
            return
        hx.check(self.lib.hx_pair_force2(C.byref(self.pot), self._p(self.pos), self._p(self.pos), self.n_own + self.n_ghost,
                                         self._p(self.nl), self.size, self.stride, hx.LAYOUT_ROW | self.lanes_flag, self.n_own, self.Lf, 1, 0.5,
                                         self._p(self.force), self._p(self.en_pot) if aux else None,
                                         self._p(self.stress) if aux else None, self.cap,
                                         self._p(self.ctl) if guarded else None, self.skin_half, self.st), "pair_force")

    def _verlet(self, finalize, integrate, guarded):
        hx = self.hx
        hx.check(self.lib.hx_verlet_finalize_integrate_range(
            self._p(self.pos), self._p(self.image), self._p(self.vel), self._p(self.force), self.n_own, self.n_own, self.cap,
            self.dt, self.Lf, int(finalize), int(integrate), self._p(self.r0) if integrate else None,
            self._p(self.ctl) if (guarded or integrate) else None, self.st), "verlet")

    def _andersen(self, guarded):
        # two-pass thermostat (same generator stream as the reference-shaped kernel, one thread per particle)
        hx = self.hx
        need = int(self.lib.hx_andersen_scratch_bytes(self.own_cap))
        if getattr(self, "andersen_scratch", None) is None or self.andersen_scratch.numel() < need:
            self.andersen_scratch = self.torch.empty(need, dtype=self.torch.uint8, device=self.vel.device)
        hx.check(self.lib.hx_andersen_finalize2(self._p(self.vel), self._p(self.force), self.n_own, self.cap, self.dt,
                                                self.sqrt_temperature, self.coll_prob, C.byref(self.rng), self._p(self.rng_state),
                                                self._p(self.andersen_scratch), need, self._p(self.ctl) if guarded else None,
                                                self.st), "andersen")

    def _ensure_force(self):
        if not self.lists_valid:
            self.update_lists()
            self.force_valid = False
        if not self.force_valid:
            self._force()
            self.force_valid = True

    def _finish_step(self):
        if self.pending_finalize:
            self._verlet(True, False, False)
            self.pending_finalize = False

    def _steps_descriptor(self):
        """hx_dd_steps for the current state (device pointers, counts, where the pushes land)."""
        hx = self.hx
        a = hx.DdSteps()
        a.position_dev[0], a.position_dev[1] = self._pos_ptr[0], self._pos_ptr[1]
        a.current = self.cur
        a.image_dev, a.velocity_dev, a.force_dev = self.image.data_ptr(), self.vel.data_ptr(), self.force.data_ptr()
        a.r0_dev = self.r0.data_ptr()
        a.n_own, a.n_ghost, a.array_size = self.n_own, self.n_ghost, self.cap
        a.timestep = self.dt
        for d in range(3):
            a.box_length[d] = self.Lf[d]
        a.nranks, a.rank = self.nranks, self.rank
        if self.nranks > 1:
            for k, dest in enumerate(("left", "right")):
                peer = self.dec.left if dest == "left" else self.dec.right
                idx = self.send_idx[dest]
                a.send_index_dev[k] = idx.data_ptr() if idx.numel() else None
                a.send_count[k] = idx.numel()
                for b in range(2):
                    # at every step all ranks write the buffer that is not their current one: my buffer b is in use when
                    # the peer's buffer b ^ peer_flip is
                    pb = b ^ self.peer_flip[dest]
                    a.peer_ghost_dev[k][b] = self._peer_pos[peer][pb] + 16 * self.peer_offset[dest]
            a.peer_inbox_dev = self.peer_inbox_table.data_ptr()
            a.inbox_dev = self._inbox_ptr
        a.timeout_cycles = self.timeout_cycles
        a.ctl_dev = self.ctl.data_ptr()
        a.skin_half = self.skin_half
        a.step_base = self.step
        a.pending_finalize = int(self.pending_finalize)
        a.potential = C.pointer(self.pot)
        if self.shells is not None:
            a.shells = C.pointer(self.shells)
        a.g_neighbour_dev = self.nl.data_ptr()
        a.size, a.stride = self.size, self.stride
        if self.kind == "andersen":
            need = int(self.lib.hx_andersen_scratch_bytes(self.own_cap))
            if getattr(self, "andersen_scratch", None) is None or self.andersen_scratch.numel() < need:
                self.andersen_scratch = self.torch.empty(need, dtype=self.torch.uint8, device=self.vel.device)
            a.rng = C.pointer(self.rng)
            a.rng_state_dev = self.rng_state.data_ptr()
            a.andersen_scratch_dev, a.andersen_scratch_bytes = self.andersen_scratch.data_ptr(), need
            a.sqrt_temperature, a.coll_prob = self.sqrt_temperature, self.coll_prob
        return a

    def run(self, steps):
        """sampler:run(steps): chunks of speculatively enqueued steps (hx_dd_enqueue_steps: one C call per chunk, the
        halo pushes, the rebuild vote and the guarded force of every step are device-side); the host looks at the
        control block once per chunk."""
        hx = self.hx
        steps = int(steps)
        if steps <= 0:
            return
        self._ensure_force()
        fuse = self.kind == "verlet"
        remaining = steps
        while remaining > 0:
            predicted = max(self.last_interval - self.since_update, 1)
            chunk = int(min(max(self.min_chunk, min(self.max_chunk, predicted + 2)), remaining))
            self.ctl[1:3].zero_()          # halted flag and step counter; the running displacement maximum stays
            desc = self._steps_descriptor()
            hx.check(self.lib.hx_dd_enqueue_steps(C.byref(desc), chunk, self.st), "hx_dd_enqueue_steps")
            ctl = self.ctl.cpu().numpy()
            self.stats["chunks"] += 1
            self.stats["syncs"] += 1
            if int(ctl[3]) == 0xDEAD:
                raise hx.HxError(f"rank {self.rank}: a peer did not reach step {self.step + int(ctl[2]) + 1} (halo wait timed out)")
            done, halted = int(ctl[2]), bool(ctl[1])
            # every executed integration kernel switched the position buffer (the halted step's ran as well)
            self.cur ^= (done + (1 if halted else 0)) & 1
            self.since_update += done
            self.step += done
            if halted:
                self.last_interval = max(self.since_update + 1, 1)
                self.update_lists()
                self._force()
                if not fuse:
                    self._andersen(False)
                done += 1
                self.step += 1
                self.since_update = 0
            elif self.since_update > self.last_interval:
                self.last_interval = self.since_update
            if fuse:
                self.pending_finalize = True
            remaining -= done
        self._finish_step()
        self.force_valid = True

    # ---- observables ------------------------------------------------------------------------------------------------------
    def thermodynamics(self):
        """observables.thermodynamics over all ranks (halmd/observables/thermodynamics.hpp:73-86)."""
        hx = self.hx
        self._finish_step()
        if not self.lists_valid:
            self.update_lists()
        self._force(aux=True)
        self.force_valid = True
        hx.check(self.lib.hx_thermodynamics(self._p(self.vel), self._p(self.en_pot), self._p(self.stress), self.cap, self.n_own,
                                            self._p(self.thermo), self.st), "thermodynamics")
        t = self.thermo[:8].clone()
        self.ex.allreduce(t, "sum")
        s = t.cpu().numpy()
        N = float(self.n_global)
        volume = float(np.prod(self.L))
        en_kin, en_pot, virial = 0.5 * s[0] / N, s[5] / N, s[6] / N
        temp = 2 * en_kin / 3
        return {"kinetic_energy": en_kin, "potential_energy": en_pot, "virial": virial, "temperature": temp,
                "pressure": N / volume * (temp + virial / 3), "internal_energy": en_kin + en_pot,
                "center_of_mass_velocity": list(s[1:4] / s[4]), "density": N / volume}

    def mean_list_length(self):
        t = self.nl_count[:self.n_own].sum(dtype=self.torch.int64).to(self.torch.float64).reshape(1)
        c = self.torch.tensor([float(self.n_own)], dtype=self.torch.float64, device=t.device)
        self.ex.allreduce(t, "sum")
        self.ex.allreduce(c, "sum")
        return float(t.item() / c.item())

    def _pinned(self, name, shape, dtype):
        buf = self._staging.get(name)
        if buf is None or buf.shape[0] < shape[0]:
            buf = self.torch.empty((int(shape[0] * 1.1) + 1024,) + tuple(shape[1:]), dtype=dtype, pin_memory=True)
            self._staging[name] = buf
        return buf[:shape[0]]

    def download(self):
        """particle:get of positions and velocities of the owned particles as double-precision host arrays: the
        double-single planes are joined on the device and leave through pinned staging buffers (48 bytes per particle)."""
        torch = self.torch
        self._finish_step()
        n, cap = self.n_own, self.cap
        out = {}
        for name, arr in (("position", self.pos), ("velocity", self.vel)):
            joined = arr[:n, :3].double() + arr[cap:cap + n, :3].double()
            host = self._pinned(name, (n, 3), torch.float64)
            host.copy_(joined, non_blocking=True)
            out[name] = host
        torch.cuda.current_stream().synchronize()
        return {k: v.numpy() for k, v in out.items()}

    def local_state(self):
        """Host copies of this rank's owned particles: ids, positions (double), velocities (double), forces, images."""
        self._ensure_force()
        n = self.n_own
        st = {k: v.copy() for k, v in self.download().items()}
        st.update({"id": self.id[:n].cpu().numpy().astype(np.int64), "force": self.force[:n, :3].cpu().numpy(),
                   "image": self.image[:n, :3].cpu().numpy()})
        return st

    def upload(self, positions, velocities):
        """particle:set of positions and velocities of the owned particles (same order as download()) from double host
        arrays, split into double-single words on the device (dsdeq, dsfun.hpp:122-126).  As in the single-GPU module
        layer, the neighbour lists stay valid unless some particle now lies further than skin / 2 from where it was at
        the last list update (the displacement test is re-evaluated over all ranks); ghosts are refreshed."""
        torch, hx, n, cap = self.torch, self.hx, self.n_own, self.cap
        self._finish_step()
        for name, arr, host in (("position", self.pos, positions), ("velocity", self.vel, velocities)):
            src = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float64))
            stage = self._staging.get(name)
            if stage is not None and src.data_ptr() == stage.data_ptr() and src.shape[0] == n:
                stage = stage[:n]                  # the caller hands back the pinned buffer download() returned
            else:
                stage = self._pinned(name + "_up", (n, 3), torch.float64)
                stage.copy_(src)
            x = stage.to(arr.device, non_blocking=True)
            hi = x.float()
            arr[:n, :3] = hi
            arr[cap:cap + n, :3] = (x - hi.double()).float()
        self.force_valid = False
        if self.lists_valid:
            self._halo_exchange(guard=False)
            hx.check(self.lib.hx_displacement_max(self._p(self.pos), self._p(self.r0), n, self.Lf, self._p(self.ctl), 1, self.st))
            self.ex.allreduce(self.ctl[0:1], "max")
            if self.lib.hx_displacement_exceeds(int(self.ctl[0].item()) & 0xFFFFFFFF, self.skin) != 0:
                self.lists_valid = False

    def neighbour_sets(self):
        """{global id: sorted global ids of its neighbours} for the owned particles (parity tests)."""
        if not self.lists_valid:
            self.update_lists()
        n, ntot = self.n_own, self.n_own + self.n_ghost
        ids = self.id[:ntot].cpu().numpy().astype(np.int64)
        rows = self.nl[:n * self.stride].cpu().numpy().view(np.uint32).reshape(n, self.stride)
        cnt = self.nl_count[:n].cpu().numpy()
        return {int(ids[i]): np.sort(ids[rows[i, :cnt[i]]]) for i in range(n)}
