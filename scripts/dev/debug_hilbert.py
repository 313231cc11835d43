import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import support
from halmd_b200 import _lib as hx
import test_gpu_reference_kernels as T
from test_gpu_kernels import Dev, lj_fluid
dev = Dev(torch)
refgpu = support.load_refgpu() if hasattr(support, "load_refgpu") else None
if refgpu is None:
    refgpu = C.CDLL(os.path.join(ROOT, "oracle/_ref/libhalmd_ref_gpu.so"))
n = 40000
L, p, v, A = lj_fluid(n, 0.8442, 62)
depth = hx.lib.hx_hilbert_depth(hx.d3(L))
p[:n, :3] = T.adversarial_positions(n, L, [1 << depth] * 3, np.random.default_rng(162))
d_pos = dev.up(p); d_keys = dev.zeros(A, np.uint32)
hx.check(hx.lib.hx_hilbert_keys(dev.p(d_pos), A, hx.f3(L), depth, dev.p(d_keys), dev.stream))
keys = dev.down(d_keys)
ref = np.zeros(A, dtype=np.uint32)
refgpu.refgpu_hilbert_map(support.ptr(p), C.c_size_t(A), support.f3(L), depth, support.ptr(ref))
bad = np.nonzero(keys != ref)[0]
print("depth", depth, "L", L, "bad", bad.size)
Lf = L.astype(np.float32)
for i in bad[:12]:
    x = p[i, :3]
    print(i, x, (x.astype(np.float64) / L + 0.5) * (1 << depth), keys[i], ref[i], np.abs(x) - Lf / 2)
