"""ctypes wrapper of the REFERENCE's own GPU step loop (oracle/ref_gpu/refsim.cpp in oracle/_ref/libhalmd_ref_gpu.so).

TEST INFRASTRUCTURE ONLY: imported by tests/ (trajectory / energy-drift comparison with the reference's kernels) and by
bench.py's `gpu_reference` leg (the recompiled-CUDA baseline timed beside the product); never by halmd_b200/.
The library is built in the development container from the reference's kernel translation units compiled where they
lie (oracle/Makefile) and travels to the GPU box as a prebuilt file; nothing here reads /root/reference at run time."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libhalmd_ref_gpu.so")


def available():
    return os.path.exists(SO)


def load():
    lib = C.CDLL(SO)
    for name in ("refsim_create", "refsim_run", "refsim_get"):
        getattr(lib, name).restype = C.c_int
    lib.refsim_destroy.restype = None
    return lib


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class ReferenceLoop:
    """One-species LJ smooth_r4 fluid stepped by the reference's kernels in the reference's call order
    (core::mdstep, halmd/mdsim/core.cpp:30-63), state resident on the device."""

    def __init__(self, hx, position, velocity, box_length, cutoff=2.5, h=0.005, skin=0.5, timestep=0.001):
        """position, velocity: (n, 3) float64 host arrays; hx: the product's ctypes module, used ONLY for the host-side
        planning formulas it shares with the reference's host classes (array size, cell grid, list size, curve depth,
        potential tables: binning.cpp:81-101, from_binning.cpp:96-119, hilbert.cpp:51-54, smooth_r4.hpp:55-99)."""
        self.lib = load()
        n = int(position.shape[0])
        L = np.asarray(box_length, dtype=np.float64)
        A = int(hx.lib.hx_array_size(n))
        self.n, self.A, self.L = n, A, L
        par = hx.LjParams()
        hx.check(hx.lib.hx_potential_lj_init(C.byref(par), 1, 1, hx.farr([1.0]), hx.farr([1.0]), hx.TRUNC_SMOOTH_R4,
                                             hx.farr([cutoff]), float(h)))
        ncell, cl, cs, size = (C.c_uint * 3)(), (C.c_float * 3)(), C.c_uint(), C.c_uint()
        hx.check(hx.lib.hx_binning_plan(hx.d3(L), n, par.r_cut[0], float(skin), 0.5, ncell, cl, C.byref(cs)))
        hx.check(hx.lib.hx_nlist_plan(par.r_cut[0], float(skin), n, float(np.prod(L)), 0.4, cs.value, C.byref(size)))
        depth = hx.lib.hx_hilbert_depth(hx.d3(L))
        pos8, vel8 = np.zeros((2 * A, 4), dtype=np.float32), np.zeros((2 * A, 4), dtype=np.float32)
        for arr, src in ((pos8, position), (vel8, velocity)):
            src = np.asarray(src, dtype=np.float64)
            hi = src.astype(np.float32)
            arr[:n, :3] = hi
            arr[A:A + n, :3] = (src - hi.astype(np.float64)).astype(np.float32)
        tags = np.zeros(A, dtype=np.uint32)
        tags[n:] = 0xFFFFFFFF                     # particle.cpp:93-113: ghost slots carry species -1
        pos8[:A, 3] = tags.view(np.float32)
        vel8[:A, 3] = 1.0
        self.h = C.c_void_p()
        rc = self.lib.refsim_create(C.c_size_t(n), C.c_size_t(A), _ptr(pos8), _ptr(vel8), _f3(L), C.c_float(par.epsilon[0]),
                                    C.c_float(par.sigma2[0]), C.c_float(par.r_cut[0]), C.c_float(par.rr_cut[0]),
                                    C.c_float(par.en_cut[0]), C.c_float(par.rri_smooth), C.c_float(skin), C.c_float(timestep), ncell,
                                    cl, cs.value, size.value, depth, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"refsim_create failed ({rc})")

    def run(self, steps):
        """Returns (device milliseconds of the loop incl. its host round trips, list rebuilds)."""
        ms, reb = C.c_float(), C.c_size_t()
        rc = self.lib.refsim_run(self.h, C.c_size_t(int(steps)), C.byref(ms), C.byref(reb))
        if rc != 0:
            raise RuntimeError(f"refsim_run failed ({rc})")
        return ms.value, reb.value

    def state(self):
        """Positions and velocities (float64 = high + low words) and forces, ordered by particle id."""
        A, n = self.A, self.n
        pos8, vel8 = np.zeros((2 * A, 4), dtype=np.float32), np.zeros((2 * A, 4), dtype=np.float32)
        img, f = np.zeros((A, 4), dtype=np.float32), np.zeros((A, 4), dtype=np.float32)
        ids = np.zeros(A, dtype=np.uint32)
        rc = self.lib.refsim_get(self.h, _ptr(pos8), _ptr(vel8), _ptr(img), _ptr(f), _ptr(ids))
        if rc != 0:
            raise RuntimeError(f"refsim_get failed ({rc})")
        order = np.argsort(ids[:n])
        pos = pos8[:n, :3].astype(np.float64) + pos8[A:A + n, :3]
        vel = vel8[:n, :3].astype(np.float64) + vel8[A:A + n, :3]
        return {"position": pos[order], "velocity": vel[order], "force": f[:n, :3][order], "image": img[:n, :3][order]}

    def close(self):
        if self.h:
            self.lib.refsim_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
