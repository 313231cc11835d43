#!/usr/bin/env python
"""The reference's own CUDA kernels (oracle/_ref/libhalmd_ref_gpu.so: its .cu translation units compiled unmodified
for sm_100a with its nvcc flags) timed on this GPU next to the product's kernels, on the same config-2-sized inputs:
pair force (smooth_r4, both list layouts), neighbour-list build (both layouts), velocity-Verlet finalize + integrate.
Device times from CUDA events around the launches; inputs are resident, caches warm for both.  Run on a GPU box:
    python scripts/ref_kernel_times.py > gpurun_out/ref_kernel_times.json"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import support  # noqa: E402
from halmd_b200 import _lib as hx  # noqa: E402
import test_gpu_kernels as tk  # noqa: E402
import test_gpu_reference_kernels as tr  # noqa: E402

torch.cuda.set_device(0)
dev = tk.Dev(torch)
oracle = support.load_oracle()
ref = C.CDLL(tr.REF_GPU_SO)
ref.refgpu_last_kernel_ms.restype = C.c_float
N = 256000
out = {"n": N, "note": "lattice + 5 % jitter at rho = 0.8442, r_c = 2.5, skin 0.5 (config 2 geometry); device ms"}


def ours(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for layout in (0, 1):
    s = tk.build_lists(hx, oracle, dev, N, 0.8442, 91, [2.5], 0.5, 0.5, 0.4, layout=layout, fill_all=1, reduce=True, jitter=0.05)
    n, A, size, stride = s["n"], s["A"], s["size"], s["stride"]
    key = "column" if layout == 0 else "row"
    out[f"mean_neighbours_{key}"] = float(s["count"].mean())
    # pair force, smooth_r4, no aux: reference compute (column lists) / compute_unroll_force_loop (row lists)
    t = []
    for _ in range(3):
        tr.ref_force(ref, hx, s, 2, [1.0], [1.0], [2.5], False, layout)
        t.append(ref.refgpu_last_kernel_ms())
    out[f"force_reference_{key}_ms"] = min(t[1:])
    par = tr.lj_params(hx, 1, [1.0], [1.0], 2, [2.5], 0.005, None)
    d_f = dev.zeros((A, 4), np.float32)
    flags = hx.FORCE_POSITIONS_REDUCED | hx.FORCE_LANES(4 if layout else 8)
    out[f"force_product_{key}_ms"] = ours(lambda: hx.check(hx.lib.hx_pair_force(
        C.byref(par), dev.p(s["d_pos"]), dev.p(s["d_pos"]), dev.p(s["d_nl"]), size, stride, layout | flags, n, hx.f3(s["L"]), 1,
        0.5, dev.p(d_f), None, None, A, dev.stream)))
    # neighbour-list build
    rr_cut_skin = np.array([(2.5 + 0.5) ** 2], dtype=np.float32)
    ref_stride = A if layout == 0 else n
    nl = np.zeros(size * ref_stride, dtype=np.uint32)
    over = C.c_int(-1)
    t = []
    for _ in range(3):
        tr.rc(ref.refgpu_nlist_build(support.ptr(s["g_cell"]), support.u3(s["ncell"]), s["cs"], support.ptr(s["p"]), C.c_size_t(A),
                                     support.farr(rr_cut_skin), 1, 1, support.f3(s["L"]), support.ptr(nl), size,
                                     C.c_size_t(ref_stride), layout, C.byref(over)), "refgpu_nlist_build")
        t.append(ref.refgpu_last_kernel_ms())
    out[f"nlist_reference_{key}_ms"] = min(t[1:])
    d_cell = dev.up(s["g_cell"])
    d_count, d_flag = dev.zeros(n, np.uint32), dev.zeros(1, np.int32)
    out[f"nlist_product_{key}_ms"] = ours(lambda: hx.check(hx.lib.hx_nlist_build(
        dev.p(d_cell), hx.u3(s["ncell"]), s["cs"], dev.p(s["d_pos"]), n, hx.farr(rr_cut_skin), 1, 1, hx.f3(s["L"]), dev.p(s["d_nl"]),
        size, stride, layout, 1, dev.p(d_count), dev.p(d_flag), dev.stream)), reps=5)

# velocity-Verlet: reference finalize + integrate (two kernels) vs the product's fused kernel
L, p, v, A = tk.lj_fluid(N, 0.8442, 92)
f = np.zeros((A, 4), dtype=np.float32)
f[:N, :3] = np.random.default_rng(92).normal(scale=10, size=(N, 3))
img = np.zeros((A, 4), dtype=np.float32)
t = []
for _ in range(3):
    rp, rv, ri = p.copy(), v.copy(), img.copy()
    tr.rc(ref.refgpu_verlet(1, 1, support.ptr(rp), support.ptr(ri), support.ptr(rv), support.ptr(f), C.c_size_t(A), C.c_float(0.001),
                            support.f3(L)), "refgpu_verlet")
    t.append(ref.refgpu_last_kernel_ms())
out["verlet_reference_finalize_plus_integrate_ms"] = min(t[1:])
d_p, d_v, d_f, d_i = dev.up(p), dev.up(v), dev.up(f), dev.up(img)
out["verlet_product_fused_ms"] = ours(lambda: hx.check(hx.lib.hx_verlet_finalize_integrate(
    dev.p(d_p), dev.p(d_i), dev.p(d_v), dev.p(d_f), N, A, 0.001, hx.f3(L), 1, None, None, dev.stream)))
print(json.dumps(out))
