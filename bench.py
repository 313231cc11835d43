#!/usr/bin/env python
"""bench.py — particle-steps/s of the MD step loop (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg1|cfg2|cfg3|cfg4|cfg5]

A bench "step" is one block of --md-steps MD steps (default 1000) of the workload, through the module API
(sampler:run).  Workload at N=1: BASELINE config 2 — 3-D LJ fluid, N=256,000, rho=0.8442, T=1.2, r_c=2.5 smooth_r4
(h=0.005), NVE velocity-Verlet dt=0.001, skin 0.5, Hilbert sort at every neighbour-list update; inputs are synthetic
(fcc lattice + Maxwell-Boltzmann velocities, melted with the Andersen thermostat before timing).

One JSON line is printed by rank 0 (see the contract in the task statement): value = whole-job particle-steps/s
with the state resident in HBM; e2e = same metric through the public API with host buffers (H2D of positions and
velocities, D2H of the results inside the timed region); roofline = pair-force kernel vs measured HBM peak;
cpu_baseline = the host-backend port (oracle/host_backend.c) timed on this box's cores on a bounded sample;
gpu_reference = the reference's OWN GPU kernels (compiled in place for sm_100a, oracle/_ref) replayed as its step loop on
the same melted state — the recompiled-CUDA baseline; detail.cfg3 / detail.cfg5 = short runs of BASELINE configs 3 and 5
as written (KA mixture with MSD + ISF blocking inside the timed region; wetting slab with walls and density modes / SSF).
--impl reference times the CPU port with all host threads as the reference arm.
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
                                                "Andersen NVT rate 2 dt=0.001, skin 0.3 occupancy 0.7, MSD + MQD + VACF + ISF "
                                                "blocking scheme (size 10, every 100 steps, semilog q-grid from 2 pi / L to 15, "
                                                "<= 7 wavevectors per shell) sampled inside the timed region"),
    "cfg5": dict(kind="slab", n=2000000, workload="wetting slab: LJ fluid N=2000000 between two planar LJ walls (slit pore, "
                                                  "rho=0.6, rc=2.5 smooth_r4), NVE dt=0.001 after an Andersen melt at T=0.9, density "
                                                  "modes along the wall normal + SSF parallel to the walls every 100 steps inside the "
                                                  "timed region"),
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
    ap.add_argument("--n-total", type=int, default=None,
                    help="total particles of the decomposed workload, divided evenly over the GPUs: strong scaling (cfg4)")
    ap.add_argument("--md-steps", type=int, default=None, help="MD steps per bench step (default 1000; 200 for cfg4)")
    ap.add_argument("--layout", default="row", choices=["row", "column"])
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the gpu_reference leg and the short cfg3 / cfg5 runs of the default line")
    ap.add_argument("--cpu-sample-steps", type=int, default=40)
    ap.add_argument("--melt-steps", type=int, default=None,
                    help="Andersen steps (dt=0.005) used to melt the lattice before timing (default 2000; 600 for cfg4)")
    a = ap.parse_args()
    if a.config is None:
        a.config = "cfg2" if a.gpus == 1 else "cfg4"
    dd = CONFIGS[a.config]["kind"] == "lj-dd"
    a.strong = bool(dd and a.n_total)
    if a.strong:
        a.n_per_gpu = a.n_total // max(a.gpus, 1)
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
class CpuHostBackend:
    """The host-backend port (oracle/host_backend.c: double precision, half lists, Newton-3) on a bounded sample of the
    workload.  The fcc lattice is melted first (untimed): velocities drawn hot enough that the fluid equilibrates near
    the target temperature, so that the timed steps see liquid-like lists and include their rebuilds.  Nothing of the
    product is imported here: the reference arm must not load libhalmd_b200.so."""

    def __init__(self, cfg, threads, seed=1, melt_steps=400, state=None):
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import support
        self.np, self.support = np, support
        oracle = self.oracle = support.load_oracle()
        oracle.host_sim_create.restype = C.c_void_p
        oracle.host_sim_rebuilds.restype = C.c_size_t
        n = self.n = cfg["n"]
        rng = np.random.default_rng(seed)
        if cfg["kind"] == "ka":
            rho, T, nt = 1.2, 0.7, 2
            eps, sig, skin = [1.0, 1.5, 1.5, 0.5], [1.0, 0.8, 0.8, 0.88], 0.3
            species = np.ascontiguousarray((np.arange(n) >= int(0.8 * n)).astype(np.uint32))
        else:
            rho, T, nt = 0.8442, 1.2, 1
            eps, sig, skin = [1.0], [1.0], 0.5
            species = None
        L = (n / rho) ** (1 / 3)
        sim = oracle.host_sim_create(C.c_size_t(n), support.d3([L] * 3), nt, support.darr(eps), support.darr(sig),
                                     support.darr([2.5] * (nt * nt)), 2, C.c_double(0.005), C.c_double(skin), C.c_double(0.001), threads)
        if not sim:
            raise RuntimeError("host_sim_create failed")
        self.sim = C.c_void_p(sim)
        if state is not None:
            pos, vel = (np.ascontiguousarray(a, dtype=np.float64) for a in state)
            self.start = "state melted on the GPU"
            melt_steps = 0
        else:
            pos = np.ascontiguousarray(support.fcc_positions(n, [L] * 3))
            # a lattice started at 1.9 T melts to about T (half of the excess kinetic energy goes into the potential)
            vel = rng.normal(scale=np.sqrt(1.9 * T), size=(n, 3))
            vel = np.ascontiguousarray(vel - vel.mean(axis=0))
            self.start = f"fcc lattice melted by {melt_steps} untimed steps from hot velocities"
        oracle.host_sim_set(self.sim, support.ptr(pos), support.ptr(vel), support.ptr(species) if species is not None else None, None)
        oracle.host_sim_run(self.sim, C.c_size_t(max(melt_steps, 2)))      # (also builds cells + lists, first forces)
        self.threads = oracle.host_sim_threads(self.sim)

    def sample(self, md_steps):
        """(particle-steps/s, seconds, list rebuilds) of md_steps MD steps continuing from the current state."""
        r0 = self.oracle.host_sim_rebuilds(self.sim)
        t0 = time.perf_counter()
        self.oracle.host_sim_run(self.sim, C.c_size_t(md_steps))
        dt = time.perf_counter() - t0
        return self.n * md_steps / dt, dt, self.oracle.host_sim_rebuilds(self.sim) - r0

    def close(self):
        self.oracle.host_sim_destroy(self.sim)


def run_reference(args, cfg):
    """Reference arm: the reference's CPU (host backend) implementation of the path — its port, since the host
    backend itself cannot be built here (Boost/Lua/HDF5 absent) — with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if cfg["kind"] == "lj-dd":
        # the CPU arm times the same fluid (bounded sample: one million particles) on the host cores
        cfg = dict(cfg, kind="lj", n=min(args.n_per_gpu or cfg["n"], 1000000))
        args.cpu_sample_steps = min(args.cpu_sample_steps, 20)
    elif cfg["kind"] == "slab":
        cfg = dict(cfg, kind="lj", n=1000000)          # (the port has no wall potential: bulk fluid of half the size)
        args.cpu_sample_steps = min(args.cpu_sample_steps, 20)
    threads = os.cpu_count() or 1
    sample = args.cpu_sample_steps
    cpu = CpuHostBackend(cfg, threads, seed=1)
    times, rebuilds = [], 0
    for i in range(args.warmup + args.steps):
        v, dt, r = cpu.sample(sample)
        if i >= args.warmup:
            times.append(dt)
            rebuilds += r
    used, start = cpu.threads, cpu.start
    cpu.close()
    value = cfg["n"] * sample * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "sample": f"{sample} MD steps of the N={cfg['n']} system per bench step, {start} "
                   f"({rebuilds} list rebuilds inside the timed steps); host backend port, OpenMP over particles = our "
                   "parallelisation (the reference host backend is single-threaded)"},
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


class Workload:
    """A wired-up single-GPU system plus what one MD block of its BASELINE config consists of."""

    def __init__(self, system, kind):
        self.s, self.kind = system, kind
        self.blocking = None          # cfg3: observables.dynamics.blocking_scheme sampled by sampler:run
        self.ssf = None               # cfg5: (ssf along x, ssf parallel to the walls), sampled every `every` steps
        self.every = 100
        self.samples = 0

    def __getattr__(self, name):
        return getattr(self.s, name)

    def run(self, steps):
        from halmd_b200 import observables
        if self.blocking is not None:
            observables.sampler.run(steps, blocking_scheme=self.blocking)
        elif self.ssf is not None:
            done = 0
            while done < steps:
                k = min(self.every, steps - done)
                self.s.run(k)
                done += k
                if k == self.every:
                    for ssf in self.ssf:
                        ssf.sample()
                    self.samples += 1
        else:
            self.s.run(steps)


def build_system(args, cfg, seed, rank=0, world=1, total_md_steps=None):
    import numpy as np
    from halmd_b200 import mdsim, observables, simulation
    n = getattr(args, "n_override", None) or cfg["n"]
    if cfg["kind"] == "lj-dd":
        n = (args.n_per_gpu or cfg["n"]) * world
        dd = simulation.decomposed_lj_fluid(n, rank, world, seed=seed, integrator="andersen", timestep=0.005, rate=2.0,
                                            lanes=args.lanes)
        dd.run(args.melt_steps)
        dd.set_integrator("verlet", timestep=0.001)
        return DecomposedAdapter(dd)
    if cfg["kind"] == "lj":
        # melt the lattice with the Andersen thermostat (examples/benchmarks/*/generate_configuration.lua), then NVE
        s = simulation.lennard_jones_fluid(n, seed=seed, layout=args.layout, integrator="andersen", timestep=0.005, rate=2.0)
        s.run(args.melt_steps)
        integ = mdsim.integrators.verlet(box=s.box, particle=s.particle, timestep=0.001)
        s.integrator, s.core = integ, mdsim.core.get()
        return Workload(s, "lj")
    if cfg["kind"] == "slab":
        s = simulation.wetting_slab(n, timestep=0.002)
        s.run(min(args.melt_steps, 600))                                   # Andersen melt at T = 0.9 between the walls
        integ = mdsim.integrators.verlet(box=s.box, particle=s.particle, timestep=0.001)
        s.integrator, s.core = integ, mdsim.core.get()
        w = Workload(s, "slab")
        Lx, Ly = s.box.length[0], s.box.length[1]

        class grid:                     # density modes along the wall normal (layering) and parallel to the walls
            def __init__(self, axis, L, kmax):
                q = np.zeros((kmax, 3))
                q[:, axis] = 2 * np.pi / L * np.arange(1, kmax + 1)
                self.value, self.shell = q, [(k, k + 1) for k in range(kmax)]
        w.ssf = [observables.ssf(density_mode=observables.density_mode(particle=s.particle, wavevector=grid(0, Lx, 96)), norm=n),
                 observables.ssf(density_mode=observables.density_mode(particle=s.particle, wavevector=grid(1, Ly, 96)), norm=n)]
        w.nq = 192
        return w
    s = simulation.kob_andersen_mixture(n, seed=seed, layout=args.layout)
    s.integrator.set_timestep(0.005)
    s.run(args.melt_steps)
    s.integrator.set_timestep(0.001)
    w = Workload(s, "ka")
    # examples/liquid/lennard_jones.lua:115-185: semilog wavenumber grid, <= 7 wavevectors per shell, blocking scheme of
    # size 10 sampled every 100 steps, maximum lag = run length / 10
    L = max(s.box.length)
    qgrid = observables.utility.semilog_grid(start=2 * np.pi / L, stop=15.0, decimation=10).value
    wv = observables.utility.wavevector(box=s.box, wavenumber=qgrid, tolerance=0.01, max_count=7)
    total = total_md_steps or (args.warmup + args.steps + 6) * args.md_steps
    w.blocking = observables.dynamics.blocking_scheme(max_lag=max(total * 0.001 / 10, 0.2), every=100, size=10, wavevector=wv)
    w.nq = len(wv.value)
    return w


def force_kernel_roofline(s, args, torch, bursts=12, per_burst=3):
    """Average duration of the pair-force kernel of the step loop, timed alone with CUDA events on its stream, L2
    flushed before every launch; algorithmic bytes = N * (32 + 4 * mean stored list length).  The
    production kernel walks the private shell rows up to a bound that grows with the displacement since the last list
    update, so its duration is averaged over states spread along rebuild intervals (4 MD steps between bursts).  For
    the decomposed workload the kernel of rank 0 is timed on its own rows (N = owned particles)."""
    from halmd_b200 import _lib as hx
    lib = hx.lib
    stream = torch.cuda.current_stream()
    st = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    kernel = "k_force_row2 (public rows, packed f32x2, 4 lanes per particle)"
    advance, kind = None, "lj"
    if isinstance(s, DecomposedAdapter):
        d = s.dd
        d._finish_step()
        if not d.lists_valid:
            d.update_lists()
        n = d.n_own
        nbar = float(d.nl_count[:n].sum().item()) / n
        force = torch.empty(d.cap * 4, dtype=torch.float32, device="cuda")

        def launch():
            if d.shells is not None:
                hx.check(lib.hx_pair_force_shells(C.byref(d.pot), d._p(d.pos), d._p(d.pos), C.byref(d.shells), d._p(d.r0), n, d.Lf,
                                                  hx.FORCE_POSITIONS_REDUCED, 1, force.data_ptr(), d._p(d.ctl), d.skin_half, st))
            else:
                hx.check(lib.hx_pair_force2(C.byref(d.pot), d._p(d.pos), d._p(d.pos), d.n_own + d.n_ghost, d._p(d.nl), d.size, d.stride,
                                            hx.LAYOUT_ROW, n, d.Lf, 1, 0.5, force.data_ptr(), None, None, d.cap, None, 0.0, st))
        if d.shells is not None and d.nranks == 1:
            kernel = "k_force_shell (private shell-ordered rows, prefix walk, packed f32x2, 4 lanes per particle)"

            def advance():
                d.run(4)
                d._finish_step()
                if not d.lists_valid:
                    d.update_lists()
    else:
        p, nb, pot = s.particle, s.neighbour, s.potential
        n = p.nparticle
        nb.check()
        nbar = nb.mean_length()
        pos = p.device_ptr("position")
        force = torch.empty(p.array_size * 4, dtype=torch.float32, device="cuda")
        L = hx.f3(s.box.length)
        par = pot.params
        skin_half = 0.25 if s.kind != "ka" else 0.15
        kind = s.kind

        def launch():
            # as the module layer launches it inside the step loop (positions reduced by the integrator); the position
            # buffers swap at list updates and inside chunks: the pointer is fetched per launch
            pos = p.device_ptr("position")
            sh = lib.hxm_neighbour_named_ptr(nb._h, b"shell_rows")
            if sh:
                hx.check(lib.hx_pair_force_shells(C.byref(par), pos, pos, C.cast(sh, C.POINTER(hx.ShellRows)),
                                                  lib.hxm_neighbour_named_ptr(nb._h, b"r0"), n, L, hx.FORCE_POSITIONS_REDUCED, 1,
                                                  force.data_ptr(), lib.hxm_neighbour_named_ptr(nb._h, b"ctl"), skin_half, st))
            else:
                hx.check(lib.hx_pair_force(C.byref(par), pos, pos, lib.hxm_neighbour_device_ptr(nb._h), nb.size, nb.stride,
                                           nb.layout | hx.FORCE_POSITIONS_REDUCED, n, L, 1, 0.5, force.data_ptr(), None, None,
                                           p.array_size, st))
        if lib.hxm_neighbour_named_ptr(nb._h, b"shell_rows"):
            kernel = "k_force_shell (private shell-ordered rows, prefix walk, packed f32x2, 4 lanes per particle)"
            def advance():
                s.s.run(4)
                nb.check()                      # (a state that is due for a list update gets it before the kernel is timed)
    times = []
    for burst in range(bursts):
        if advance is not None and burst > 0:
            advance()
        for i in range(per_burst + (1 if burst == 0 else 0)):
            flush.fill_(float(i))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            launch()
            e1.record(stream)
            torch.cuda.synchronize()
            if not (burst == 0 and i == 0):
                times.append(e0.elapsed_time(e1) * 1e-3)
    t = sum(times) / len(times)
    bytes_alg = n * (32.0 + 4.0 * nbar)
    peak, how = measured_peaks()
    achieved = bytes_alg / t / 1e9
    # secondary roofline: FP32 lane operations of the packed instruction stream actually needed — 41 per list entry
    # INSIDE the cutoff (warps that skip the minimum image; 50 otherwise) plus 7 per stored entry for the distance
    # test — against 128 lanes x SM count x SM clock
    fp32_peak = 148 * 128 * 1.965e9
    inside = nbar * (2.5 / 3.0) ** 3 if kind != "ka" else nbar * (2.5 / 2.8) ** 3
    # DRAM bytes of one launch from the committed ncu --set full capture of the same kernel and workload, scaled to n
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_force_traffic.json")
    if os.path.exists(tpath) and not isinstance(s, DecomposedAdapter) and kind == "lj":
        tj = json.load(open(tpath))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * (n / tj["particles"])
        traffic_src = "static: ncu --set full capture in profiles/r2_force_traffic.json (one launch mid-interval), scaled to N"
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": how, "avg_launch_us": t * 1e6, "min_launch_us": min(times) * 1e6, "max_launch_us": max(times) * 1e6,
            "algorithmic_bytes_per_launch": bytes_alg, "mean_list_length": nbar, "pairs_per_s": n * nbar / t,
            "fp32_lane_ops_frac": n * (inside * 41.0 + nbar * 7.0) / t / fp32_peak,
            "timing": f"{len(times)} launches in {bursts} bursts spread over rebuild intervals, L2 flushed (256 MiB write) before each"}


def gpu_reference_leg(s, torch, steps=200):
    """The reference's own GPU kernels (oracle/_ref/libhalmd_ref_gpu.so: its .cu files compiled in place for sm_100a with
    its flags) replayed as its step loop — per-launch synchronisation, 64-float read-back every step, radix sorts and
    default column lists included — from the current melted state of the product.  A reported baseline, never part
    of `value`."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import refsim
    from halmd_b200 import _lib as hx
    if not refsim.available():
        return {"unavailable": "oracle/_ref/libhalmd_ref_gpu.so not built"}
    p = s.particle
    order = np.argsort(p.get("id"))
    ref = refsim.ReferenceLoop(hx, p.get("position_ds")[order], p.get("velocity_ds")[order], s.box.length)
    ref.run(100)                          # warm-up: clocks, first list updates
    # the loop synchronises with the host after every launch, so its time depends on host-side latencies that vary from
    # box to box and run to run (439, 783 and 4304 us per step were seen for the same state): the fastest of several
    # segments is reported — the baseline at its best
    segments = []
    for _ in range(4):
        ms, rebuilds = ref.run(steps // 2)
        segments.append((1e3 * ms / (steps // 2), int(rebuilds)))
    ref.close()
    n = p.nparticle
    best, rebuilds = min(segments)
    steps = steps // 2
    ms = best * steps * 1e-3
    return {"value": n * steps / (ms * 1e-3), "unit": "particle-steps/s", "us_per_step": 1e3 * ms / steps, "md_steps": steps,
            "segments_us_per_step": [round(x, 1) for x, _ in segments],
            "list_rebuilds": int(rebuilds), "kind": "reference kernels (recompiled for sm_100a) in the reference's call order",
            "what": "verlet integrate, max_displacement + 64-float D2H, [hilbert map, radix sort, rearrange, binning, from_binning], "
                    "pair_trunc compute (column lists), verlet finalize; cudaDeviceSynchronize after every launch"}


def extra_config_leg(args, name, md_steps, torch):
    """Short run of another BASELINE config as written (detail of the default line)."""
    from halmd_b200 import _lib as hx
    cfg = CONFIGS[name]
    a = argparse.Namespace(**vars(args))
    a.md_steps, a.melt_steps, a.warmup, a.steps = md_steps, (1000 if name == "cfg3" else 500), 1, 2
    w = build_system(a, cfg, seed=1, total_md_steps=4 * md_steps)
    w.run(md_steps)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    launches0 = hx.lib.hx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    w.run(2 * md_steps)
    e1.record(stream)
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    out = {"workload": cfg["workload"], "value": cfg["n"] * 2 * md_steps / t, "unit": "particle-steps/s", "md_steps": 2 * md_steps,
           "us_per_md_step": 1e6 * t / (2 * md_steps), "wavevectors": getattr(w, "nq", 0),
           "gpu_launches": int(hx.lib.hx_launch_count() - launches0), "temperature": w.thermodynamics()["temperature"]}
    if w.blocking is not None:
        msd = w.blocking.mean_square_displacement()
        out["msd_blocks_accumulated"] = int((msd["count"] > 0).sum())
    return out


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
    single = not decomposed
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
        st0 = s.dd.download()
        pos_h, vel_h = st0["position"], st0["velocity"]
    else:
        # host buffers of the end-to-end leg in pinned memory (the samples a caller keeps between blocks)
        pos_h = s.particle.get("position_ds", out=torch.empty(s.particle.nparticle, 3, dtype=torch.float64).pin_memory().numpy())
        vel_h = s.particle.get("velocity_ds", out=torch.empty(s.particle.nparticle, 3, dtype=torch.float64).pin_memory().numpy())
    h2d = pos_h.nbytes + vel_h.nbytes
    d2h = pos_h.nbytes + vel_h.nbytes + 10 * 8
    e2e_times = []
    for i in range(2 + min(args.steps, 3)):
        barrier()
        t0 = time.perf_counter()
        if decomposed:
            s.dd.upload(pos_h, vel_h)
            s.run(md)
            st1 = s.dd.download()
            # (ownership may have changed during the block: the next upload uses the new arrays)
            pos_h, vel_h = st1["position"], st1["velocity"]
        else:
            s.particle.set("position_ds", pos_h)
            s.particle.set("velocity_ds", vel_h)
            s.run(md)
            s.particle.get("position_ds", out=pos_h)
            s.particle.get("velocity_ds", out=vel_h)
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

    dd_stats = dict(s.dd.stats, bytes_sent_rank0=int(s.dd.ex.bytes_sent)) if decomposed else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    core_stats = s.core.stats() if single else None
    one_rank = None
    if decomposed and world > 1 and not args.no_extra:
        # the same decomposed code on one rank with the same number of particles per GPU: the denominator of the
        # weak-scaling efficiency (rank 0 only; the other ranks have left)
        del s
        torch.cuda.empty_cache()
        from halmd_b200 import simulation
        dd1 = simulation.decomposed_lj_fluid(n_rank, 0, 1, seed=1, integrator="andersen", timestep=0.005, rate=2.0)
        dd1.run(min(args.melt_steps, 300))
        dd1.set_integrator("verlet", timestep=0.001)
        dd1.run(md)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        dd1.run(2 * md)
        f1.record(stream)
        torch.cuda.synchronize()
        t1 = f0.elapsed_time(f1) * 1e-3
        one_rank = {"value": n_rank * 2 * md / t1, "unit": "particle-steps/s", "particles": n_rank, "md_steps": 2 * md, "stats": dict(dd1.stats), "updates": dd1.updates,
                    "what": "the decomposed step loop on ONE rank (no neighbours) with the same particles per GPU"}
        s = DecomposedAdapter(dd1)
    roof = force_kernel_roofline(s, args, torch)
    cpu = gpu_ref = None
    extra = {}
    if world == 1 and single and cfg["kind"] in ("lj", "ka"):
        order = np.argsort(s.particle.get("id"))
        state = (s.particle.get("position_ds")[order], s.particle.get("velocity_ds")[order])
        # (the GPU leg first, while the clocks are up: its per-launch synchronisation keeps the GPU lightly loaded, and
        # after ten idle seconds of the CPU leg it was once measured ten times slower)
        if not args.no_extra and cfg["kind"] == "lj":
            ref_clocks = ClockSampler(local)
            ref_clocks.start()
            gpu_ref = gpu_reference_leg(s, torch)
            gpu_ref["clocks"] = ref_clocks.stop()
        if not args.no_cpu_baseline:
            host = CpuHostBackend(cfg, 1, state=state)
            v, dt, reb = host.sample(args.cpu_sample_steps)
            cpu = {"value": v, "unit": "particle-steps/s", "cores": host.threads, "kind": "port",
                   "sample": f"host-backend port (double, half lists, Newton-3), N={cfg['n']}, {args.cpu_sample_steps} MD steps from the "
                             f"state melted on the GPU ({reb} list rebuilds inside), {dt:.1f} s"}
            host.close()
    if world == 1 and single and args.config == "cfg2" and not args.no_extra:
        del s
        torch.cuda.empty_cache()
        for name, md_extra in (("cfg3", 500), ("cfg5", 300)):
            try:
                extra[name] = extra_config_leg(args, name, md_extra, torch)
            except Exception as exc:      # the headline line must not be lost to a side leg
                extra[name] = {"error": repr(exc)}
            torch.cuda.empty_cache()
    e_drift = abs(th1["internal_energy"] - th0["internal_energy"]) / abs(th0["internal_energy"])
    if decomposed:
        par = (f"slab domain decomposition along x over {world} GPU(s), {n_rank} particles per GPU; every step the boundary particles' "
               "positions are stored into the neighbours' ghost slots over peer memory (NVLink) and the rebuild vote travels "
               "through per-rank inboxes, device-side wait, no host or library collective inside a step; NCCL only at list updates")
        core = dd_stats
    else:
        par = "single GPU" if world == 1 else f"{world} independent replicas (one per GPU)"
        core = core_stats
    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "strong" if getattr(args, "strong", False) else "weak",
        "vs_baseline": None, "dtype": "f32 pair math, double-single (2xf32) state, compensated f32 force sums", "data": "synthetic",
        "config": {"workload": cfg["workload"], "md_steps_per_bench_step": md, "neighbour_layout": args.layout, "lanes_per_particle": args.lanes,
                   "particles_total": world * n_rank, "parallelism": par,
                   "l2": ("step loop: the working set of a step (walked list prefixes + state, about 90 MB at N=256000) fits the 126 MB "
                          "L2 and is not flushed; roofline leg: L2 flushed before every launch") if single else
                         "step loop: working set far larger than L2"},
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
        "cpu_baseline": cpu,
        "gpu_reference": gpu_ref,
        "detail": {"ns_per_particle_step": 1e9 * elapsed / (world * n_rank * md * args.steps), "list_updates": int(nupdates),
                   "md_steps": md * args.steps, "steps_per_update": md * args.steps / max(nupdates, 1),
                   "energy_drift_rel": e_drift, "temperature": th1["temperature"], "core": core, "per_gpu_1rank": one_rank,
                   "per_gpu": value / world, **extra},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    cfg = CONFIGS[args.config]
    if args.strong:
        cfg = dict(cfg, workload=cfg["workload"].replace(
            "N=8,000,000 per GPU (64M on 8 GPUs)", f"N={args.n_per_gpu * args.gpus:,} in total, {args.n_per_gpu:,} per GPU (strong scaling)"))
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
