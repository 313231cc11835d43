#!/usr/bin/env python
"""bench.py — particle-steps/s of the MD step loop (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg1]

A bench "step" is one block of --md-steps MD steps (default 1000) of the workload, through the module API
(sampler:run).  Workload at N=1: BASELINE config 2 — 3-D LJ fluid, N=256,000, rho=0.8442, T=1.2, r_c=2.5 smooth_r4
(h=0.005), NVE velocity-Verlet dt=0.001, skin 0.5, Hilbert sort at every neighbour-list update; inputs are synthetic
(fcc lattice + Maxwell-Boltzmann velocities, melted with the Andersen thermostat before timing).

One JSON line is printed by rank 0 (see the contract in the task statement): value = whole-job particle-steps/s
with the state resident in HBM; e2e = same metric through the public API with host buffers (H2D of positions and
velocities, D2H of the results inside the timed region); roofline = pair-force kernel vs measured HBM peak;
cpu_baseline = the host-backend port (oracle/host_backend.c) timed on this box's cores on a bounded sample.
--impl reference times that CPU port with all host threads as the reference arm.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg1": dict(kind="lj", n=4000, workload="3D LJ fluid N=4000 rho=0.8442 T=1.2 rc=2.5 smooth_r4 NVE dt=0.001"),
    "cfg2": dict(kind="lj", n=256000, workload="3D LJ fluid N=256000 rho=0.8442 T=1.2 rc=2.5 smooth_r4 h=0.005 NVE dt=0.001, "
                                               "skin 0.5, Hilbert sort at every neighbour-list update"),
    "cfg4": dict(kind="lj-dd", n=8000000, workload="3D LJ fluid N=8,000,000 per GPU (64M on 8 GPUs) rho=0.8442 T=1.2 rc=2.5 smooth_r4 "
                                                  "h=0.005 NVE dt=0.001, skin 0.5, slab domain decomposition along x with ghost-layer exchange, "
                                                  "particle migration and Hilbert sort at every neighbour-list update"),
    "cfg3": dict(kind="ka", n=1000000, workload="Kob-Andersen 80:20 N=1000000 rho=1.2 T=0.7 rc=2.5sigma smooth_r4, "
                                                "Andersen NVT rate 2 dt=0.001, skin 0.3 occupancy 0.7"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS),
                    help="default: cfg2 on one GPU, cfg4 (domain-decomposed, weak scaling) on several")
    ap.add_argument("--n-per-gpu", type=int, default=None, help="particles per GPU of the decomposed workload (cfg4)")
    ap.add_argument("--md-steps", type=int, default=None, help="MD steps per bench step (default 1000; 200 for cfg4)")
    ap.add_argument("--layout", default="row", choices=["row", "column"])
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-steps", type=int, default=40)
    ap.add_argument("--melt-steps", type=int, default=None,
                    help="Andersen steps (dt=0.005) used to melt the lattice before timing (default 2000; 600 for cfg4)")
    a = ap.parse_args()
    if a.config is None:
        a.config = "cfg2" if a.gpus == 1 else "cfg4"
    dd = CONFIGS[a.config]["kind"] == "lj-dd"
    if a.md_steps is None:
        a.md_steps = 200 if dd else 1000
    if a.melt_steps is None:
        a.melt_steps = 600 if dd else 2000
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.stop_flag, self.index = [], threading.Event(), index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def start(self):
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        self.thread.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        smax = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU legs (the only places allowed to execute oracle/) ----------------------------------------------------------
def cpu_host_backend(cfg, md_steps, threads, seed=1):
    """Time the host-backend port (double, half lists, Newton-3) on a bounded sample of the workload."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import support
    from halmd_b200 import simulation
    oracle = support.load_oracle()
    n = cfg["n"]
    rng = np.random.default_rng(seed)
    if cfg["kind"] == "lj":
        rho, T, nt = 0.8442, 1.2, 1
        eps, sig, skin = [1.0], [1.0], 0.5
        species = None
    else:
        rho, T, nt = 1.2, 0.7, 2
        eps, sig, skin = [1.0, 1.5, 1.5, 0.5], [1.0, 0.8, 0.8, 0.88], 0.3
        species = np.ascontiguousarray((np.arange(n) >= int(0.8 * n)).astype(np.uint32))
    L = (n / rho) ** (1 / 3)
    pos = np.ascontiguousarray(simulation.fcc_positions(n, [L] * 3))
    vel = np.ascontiguousarray(simulation.boltzmann_velocities(n, T, rng))
    sim = oracle.host_sim_create(C.c_size_t(n), support.d3([L] * 3), nt, support.darr(eps), support.darr(sig),
                                 support.darr([2.5] * (nt * nt)), 2, C.c_double(0.005), C.c_double(skin), C.c_double(0.001), threads)
    if not sim:
        raise RuntimeError("host_sim_create failed")
    sim = C.c_void_p(sim)
    oracle.host_sim_set(sim, support.ptr(pos), support.ptr(vel), support.ptr(species) if species is not None else None, None)
    oracle.host_sim_run(sim, C.c_size_t(2))          # builds cells + lists, first forces (untimed)
    t0 = time.perf_counter()
    oracle.host_sim_run(sim, C.c_size_t(md_steps))
    dt = time.perf_counter() - t0
    used = oracle.host_sim_threads(sim)
    rebuilds = oracle.host_sim_rebuilds(sim)
    oracle.host_sim_destroy(sim)
    return n * md_steps / dt, dt, used, rebuilds


def run_reference(args, cfg):
    if cfg["kind"] == "lj-dd":
        # the CPU arm times the same fluid (per-GPU share of the decomposed workload) on the host cores
        # (bounded sample: one million particles of the same fluid)
        cfg = dict(cfg, kind="lj", n=min(args.n_per_gpu or cfg["n"], 1000000))
        args.cpu_sample_steps = min(args.cpu_sample_steps, 20)
    """Reference arm: the reference's CPU (host backend) implementation of the path — its port, since the host
    backend itself cannot be built here (Boost/Lua/HDF5 absent) — with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.cpu_sample_steps
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, dt, used, _ = cpu_host_backend(cfg, sample, threads, seed=1 + i)
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = cfg["n"] * sample * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "sample": f"{sample} MD steps of the full N={cfg['n']} system per bench step "
                   "(host backend port, OpenMP over particles = our parallelisation; the reference host backend is single-threaded)"},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": used, "kind": "port",
                         "sample": f"N={cfg['n']}, {sample} MD steps x {args.steps}"},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- our arm ----------------------------------------------------------------------------------------------------------------
class DecomposedAdapter:
    """The decomposed system behind the few names the bench uses (run / thermodynamics / counters)."""

    def __init__(self, dd):
        self.dd = dd

    def run(self, steps):
        self.dd.run(steps)

    def thermodynamics(self):
        return self.dd.thermodynamics()

    @property
    def updates(self):
        return self.dd.updates


def build_system(args, cfg, seed, rank=0, world=1):
    from halmd_b200 import mdsim, simulation
    if cfg["kind"] == "lj-dd":
        n = (args.n_per_gpu or cfg["n"]) * world
        dd = simulation.decomposed_lj_fluid(n, rank, world, seed=seed, integrator="andersen", timestep=0.005, rate=2.0,
                                            lanes=args.lanes)
        dd.run(args.melt_steps)
        dd.set_integrator("verlet", timestep=0.001)
        return DecomposedAdapter(dd)
    if cfg["kind"] == "lj":
        # melt the lattice with the Andersen thermostat (examples/benchmarks/*/generate_configuration.lua), then NVE
        s = simulation.lennard_jones_fluid(cfg["n"], seed=seed, layout=args.layout, integrator="andersen", timestep=0.005, rate=2.0)
        s.run(args.melt_steps)
        integ = mdsim.integrators.verlet(box=s.box, particle=s.particle, timestep=0.001)
        s.integrator, s.core = integ, mdsim.core.get()
        return s
    s = simulation.kob_andersen_mixture(cfg["n"], seed=seed, layout=args.layout)
    s.integrator.set_timestep(0.005)
    s.run(args.melt_steps)
    s.integrator.set_timestep(0.001)
    return s


def force_kernel_roofline(s, args, torch, iters=20):
    """Average duration of the pair-force kernel, timed alone with CUDA events on its stream, L2 flushed between
    launches; algorithmic bytes = N * (32 + 4 * mean list length).  For the decomposed workload the
    kernel of rank 0 is timed on its own rows (N = owned particles)."""
    from halmd_b200 import _lib as hx
    lib = hx.lib
    stream = torch.cuda.current_stream()
    st = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    if isinstance(s, DecomposedAdapter):
        d = s.dd
        d._finish_step()
        if not d.lists_valid:
            d.update_lists()
        n = d.n_own
        nbar = float(d.nl_count[:n].sum().item()) / n
        force = torch.empty(d.cap * 4, dtype=torch.float32, device="cuda")

        def launch():
            hx.check(lib.hx_pair_force2(C.byref(d.pot), d._p(d.pos), d._p(d.pos), d.n_own + d.n_ghost, d._p(d.nl), d.size, d.stride,
                                        hx.LAYOUT_ROW, n, d.Lf, 1, 0.5, force.data_ptr(), None, None, d.cap, None, 0.0, st))
    else:
        p, nb, pot = s.particle, s.neighbour, s.potential
        n = p.nparticle
        nb.check()
        nbar = nb.mean_length()
        pos = p.device_ptr("position")
        force = torch.empty(p.array_size * 4, dtype=torch.float32, device="cuda")
        L = hx.f3(s.box.length)
        par = pot.params

        def launch():
            # positions have been through the integrator: same flag as the module layer passes in the step loop
            hx.check(lib.hx_pair_force(C.byref(par), pos, pos, lib.hxm_neighbour_device_ptr(nb._h), nb.size, nb.stride,
                                       nb.layout | hx.FORCE_POSITIONS_REDUCED, n,
                                       L, 1, 0.5, force.data_ptr(), None, None, p.array_size, st))
    times = []
    for i in range(iters + 3):
        flush.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launch()
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1) * 1e-3)
    t = sum(times) / len(times)
    bytes_alg = n * (32.0 + 4.0 * nbar)
    peak, how = measured_peaks()
    achieved = bytes_alg / t / 1e9
    # secondary roofline: FP32 lane operations of the packed instruction stream (41 per evaluated list entry in warps
    # that skip the minimum image, 50 otherwise; the lower count is used: DESIGN.md "pair force") against
    # 128 lanes x SM count x SM clock
    fp32_peak = 148 * 128 * 1.965e9
    # DRAM bytes of one launch from the committed ncu --set full capture (same kernel, same workload), scaled to n
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_force_traffic.json")
    if os.path.exists(tpath) and not isinstance(s, DecomposedAdapter):
        tj = json.load(open(tpath))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * (n / tj["particles"])
    return {"bound": "hbm", "kernel": "k_force_row2 (packed f32x2, %d lanes per particle)" % args.lanes, "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": how, "avg_launch_us": t * 1e6,
            "algorithmic_bytes_per_launch": bytes_alg, "mean_list_length": nbar, "pairs_per_s": n * nbar / t,
            "fp32_lane_ops_frac": n * nbar * 41.0 / t / fp32_peak,
            "l2": "flushed between launches (256 MiB write)"}


def run_ours(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist
    from halmd_b200 import _lib as hx

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — halmd_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    decomposed = cfg["kind"] == "lj-dd"

    # one GPU: the whole system; several GPUs: one slab of the decomposed system per rank (weak scaling: the
    # particle number per GPU is fixed), or independent replicas for the single-GPU configurations
    s = build_system(args, cfg, seed=1 if decomposed else 1 + rank, rank=rank, world=world)
    n_rank = (args.n_per_gpu or cfg["n"]) if decomposed else cfg["n"]
    md = args.md_steps
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def updates():
        return s.updates if decomposed else s.neighbour.updates

    for _ in range(args.warmup):
        s.run(md)
    barrier()
    th0 = s.thermodynamics()
    updates0, launches0 = updates(), hx.lib.hx_launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        s.run(md)
    e1.record(stream)
    barrier()
    elapsed = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    clk = clocks.stop() if rank == 0 else None
    launches = hx.lib.hx_launch_count() - launches0
    nupdates = updates() - updates0
    th1 = s.thermodynamics()
    value = world * n_rank * md * args.steps / elapsed

    # end to end through the public API with host buffers: H2D of the state, block of MD steps, D2H of the results
    if decomposed:
        st0 = s.dd.local_state()
        pos_h, vel_h = st0["position"], st0["velocity"]
    else:
        pos_h = s.particle.get("position_ds")
        vel_h = s.particle.get("velocity_ds")
    h2d = pos_h.nbytes + vel_h.nbytes
    d2h = pos_h.nbytes + vel_h.nbytes + 10 * 8
    e2e_times = []
    for i in range(2 + min(args.steps, 3)):
        barrier()
        t0 = time.perf_counter()
        if decomposed:
            s.dd.upload(pos_h, vel_h)
            s.run(md)
            st1 = s.dd.local_state()
            # (ownership may have changed during the block: the next upload uses the new arrays)
            pos_h, vel_h = st1["position"], st1["velocity"]
        else:
            s.particle.set("position_ds", pos_h)
            s.particle.set("velocity_ds", vel_h)
            s.run(md)
            pos_h = s.particle.get("position_ds")
            vel_h = s.particle.get("velocity_ds")
        th = s.thermodynamics()
        torch.cuda.synchronize()
        if i >= 2:
            e2e_times.append(time.perf_counter() - t0)
    e2e_t = max_over_ranks(sum(e2e_times) / len(e2e_times))
    e2e_value = world * n_rank * md / e2e_t
    if world > 1:
        t = torch.tensor([float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h2d, d2h = int(t[0].item()), int(t[1].item())
        halo_bytes = torch.tensor([float(s.dd.ex.bytes_sent)], device="cuda", dtype=torch.float64) if decomposed else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    roof = force_kernel_roofline(s, args, torch)
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not decomposed:
        v, dt, used, _ = cpu_host_backend(cfg, args.cpu_sample_steps, 1)
        cpu = {"value": v, "unit": "particle-steps/s", "cores": used, "kind": "port",
               "sample": f"host-backend port (double, half lists), N={cfg['n']}, {args.cpu_sample_steps} MD steps from the melted-lattice start, {dt:.1f} s"}
    e_drift = abs(th1["internal_energy"] - th0["internal_energy"]) / abs(th0["internal_energy"])
    if decomposed:
        par = f"slab domain decomposition along x over {world} GPU(s), {n_rank} particles per GPU, ghost exchange + rebuild vote every step (NCCL)"
        core = dict(s.dd.stats, bytes_sent_rank0=int(s.dd.ex.bytes_sent))
    else:
        par = "single GPU" if world == 1 else f"{world} independent replicas (one per GPU)"
        core = s.core.stats()
    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 pair math, double-single (2xf32) state, compensated f32 force sums", "data": "synthetic",
        "config": {"workload": cfg["workload"], "md_steps_per_bench_step": md, "neighbour_layout": args.layout, "lanes_per_particle": args.lanes,
                   "particles_total": world * n_rank, "parallelism": par,
                   "l2": "working set (neighbour lists + state) exceeds the 126 MB L2; no flush inside the step loop"},
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
        "cpu_baseline": cpu,
        "detail": {"ns_per_particle_step": 1e9 * elapsed / (world * n_rank * md * args.steps), "list_updates": int(nupdates),
                   "md_steps": md * args.steps, "steps_per_update": md * args.steps / max(nupdates, 1),
                   "energy_drift_rel": e_drift, "temperature": th1["temperature"], "core": core},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
