"""Test support: ctypes access to the CPU oracle (oracle/_build/liboracle.so), to the reference's own
header-only numerics (oracle/_ref/libhalmd_ref.so, optional) and synthetic input generators.

The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libhalmd_ref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

PH = 0xFFFFFFFF
MAXPT = 16


def build_oracle():
    """Compile the C oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "--no-print-directory"], check=True, capture_output=True)


def load_oracle():
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    lib = C.CDLL(ORACLE_SO)
    lib.oracle_array_size.restype = C.c_size_t
    lib.oracle_array_size.argtypes = [C.c_size_t]
    lib.oracle_nlist_size.restype = C.c_uint
    lib.oracle_nlist_size.argtypes = [C.c_float, C.c_float, C.c_size_t, C.c_double, C.c_double, C.c_uint]
    lib.oracle_hilbert_depth.restype = C.c_uint
    lib.oracle_hilbert_map_unit.restype = C.c_uint32
    lib.oracle_hilbert_map_unit.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint]
    lib.oracle_displacement_max.restype = C.c_float
    lib.oracle_binning_update.restype = C.c_int
    lib.oracle_nlist_build.restype = C.c_int
    lib.oracle_pair_f.restype = C.c_int
    lib.oracle_pair_f.argtypes = [C.c_void_p, C.c_uint, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.oracle_pair_d.argtypes = [C.c_double] * 3 + [C.c_int, C.c_double, C.c_double,
                                                    C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.host_sim_create.restype = C.c_void_p
    lib.host_sim_rebuilds.restype = C.c_size_t
    lib.host_sim_pairs.restype = C.c_size_t
    lib.host_sim_threads.restype = C.c_int
    return lib


def load_ref():
    """Reference header shim, or None when it was not built (no /root/reference at build time)."""
    if not os.path.exists(REF_SO):
        return None
    lib = C.CDLL(REF_SO)
    lib.ref_hilbert_map.restype = C.c_uint
    lib.ref_hilbert_map.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint]
    return lib


class OLjParams(C.Structure):
    _fields_ = [
        ("ntype1", C.c_uint), ("ntype2", C.c_uint), ("truncation", C.c_int), ("rri_smooth", C.c_float),
        ("epsilon", C.c_float * MAXPT), ("sigma2", C.c_float * MAXPT), ("r_cut", C.c_float * MAXPT),
        ("rr_cut", C.c_float * MAXPT), ("en_cut", C.c_float * MAXPT), ("force_cut", C.c_float * MAXPT),
        ("hard_core", C.c_int), ("r_core", C.c_float * MAXPT),
    ]


class OPlanarWall(C.Structure):
    _fields_ = [
        ("nwall", C.c_uint), ("nspecies", C.c_uint), ("smoothing", C.c_float),
        ("surface_normal", (C.c_float * 3) * 8), ("offset", C.c_float * 8),
        ("epsilon", C.c_float * 32), ("sigma", C.c_float * 32), ("wetting", C.c_float * 32), ("cutoff", C.c_float * 32),
    ]


# parameters and test positions of the reference's unit test (test/unit/mdsim/potentials/external/planar_wall.cpp:85-147)
REF_WALL = dict(
    offset=[-5.0, 5.0, 10.0], surface_normal=[[1, 0, 0], [1, -1, 0], [-1, 0, 2]],
    cutoff=[[1.0, 2.0], [10.0, 10.0], [1e2, 1e2]], epsilon=[[1.0, 0.5], [2.0, 0.0], [0.0, 1.0]],
    sigma=[[1.0, 2.0], [3.0, 1.0], [2.0, 1.0]], wetting=[[0.4, 0.0], [-1.0, 1.0], [0.0, 2.0]], smoothing=0.01,
)
REF_WALL_POSITIONS = [[0, 0, 0], [5, 0, 1], [-4, 0, 2], [-5.7, 0, 3], [0, -3, 4], [-5.2, -8, 5], [0, 0, 20]]


def make_oracle_wall(offset, surface_normal, epsilon, sigma, wetting, cutoff, smoothing):
    """OPlanarWall filled like planar_wall.cpp:83-112 (normalised normals, tables [species * nwall + wall])."""
    w = OPlanarWall()
    eps = np.asarray(epsilon, dtype=np.float32)
    nwall, nspecies = eps.shape
    w.nwall, w.nspecies, w.smoothing = nwall, nspecies, smoothing
    for i in range(nwall):
        n = np.asarray(surface_normal[i], dtype=np.float32)
        n = n / np.sqrt(np.float32(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]))
        for d in range(3):
            w.surface_normal[i][d] = n[d]
        w.offset[i] = offset[i]
        for j in range(nspecies):
            w.epsilon[j * nwall + i] = epsilon[i][j]
            w.sigma[j * nwall + i] = sigma[i][j]
            w.wetting[j * nwall + i] = wetting[i][j]
            w.cutoff[j * nwall + i] = cutoff[i][j]
    return w


def reference_test_wall(r, normal, offset, epsilon, sigma, wetting, cutoff, smoothing):
    """compute_wall of the reference's unit test (planar_wall.cpp:48-80), in double precision."""
    normal = np.asarray(normal, dtype=np.float64)
    d = float(np.dot(r, normal) - offset)
    sign = 1.0 if d > 0 else -1.0
    d = abs(d)
    if d > cutoff:
        return np.zeros(3), 0.0
    fval = 3 * sign * epsilon / sigma * ((2. / 5) * (sigma / d) ** 10 - wetting * (sigma / d) ** 4)
    en = epsilon * ((2. / 15) * (sigma / d) ** 9 - wetting * (sigma / d) ** 3)
    en -= epsilon * ((2. / 15) * (sigma / cutoff) ** 9 - wetting * (sigma / cutoff) ** 3)
    xi = (d - cutoff) / smoothing
    w = xi ** 4 / (1 + xi ** 4)
    w1 = 4 * xi ** 3 / (1 + xi ** 4) ** 2
    fval = fval * w - w1 * en / smoothing
    en *= w
    return fval * normal, en


class ORand48(C.Structure):
    _fields_ = [("blocks", C.c_uint), ("threads", C.c_uint), ("a", C.c_uint16 * 3), ("c", C.c_uint16 * 3)]


def ptr(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def u3(v):
    return (C.c_uint * 3)(*[int(x) for x in v])


def farr(v):
    v = [float(x) for x in np.asarray(v).ravel()]
    return (C.c_float * len(v))(*v)


def darr(v):
    v = [float(x) for x in np.asarray(v).ravel()]
    return (C.c_double * len(v))(*v)


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


# ---- synthetic inputs ------------------------------------------------------------------------------

def array_size(n):
    return ((n + 1023) // 1024 if n > 0 else 1) * 1024


def fcc_positions(n, L):
    """fcc lattice filling the box (halmd/mdsim/gpu/positions/lattice.cpp:116-177,
    halmd/mdsim/positions/lattice_primitive.hpp:121-129), centred on the origin."""
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
    N = int(u * np.prod(ncell))
    skip = int(np.ceil(N / (N - n))) if N > n else 0
    idx = np.arange(N, dtype=np.int64)
    if skip:
        keep = np.ones(N, dtype=bool)
        keep[skip - 1::skip] = False
        idx = idx[keep][:n]
    idx = idx[:n]
    cell = idx >> 2
    cx = cell % ncell[0]
    cy = (cell // ncell[0]) % ncell[1]
    cz = cell // (ncell[0] * ncell[1])
    r = np.empty((n, 3), dtype=np.float64)
    r[:, 0] = cx + (((idx ^ (idx >> 1)) & 1) * 2 + 1) / 4.0
    r[:, 1] = cy + ((idx & 1) * 2 + 1) / 4.0
    r[:, 2] = cz + ((idx & 2) + 1) / 4.0
    r *= a
    r += (L - ncell * a) / 2      # centre the lattice in the box
    r -= L / 2                    # box spans (-L/2, L/2)
    return r


_RCP_FN = C.CFUNCTYPE(C.c_float, C.c_float)
_rcp_callbacks = []


def install_device_rcp(hx, oracle):
    """GPU runs: the oracle's restatement of the reference build's approximate division (r * MUFU.RCP(l) for cell and
    box lengths) asks the device for the hardware reciprocal of those constants through the product's hx_device_rcp."""
    def rcp(x):
        xin, out = (C.c_float * 1)(x), (C.c_float * 1)()
        hx.check(hx.lib.hx_device_rcp(xin, out, 1), "hx_device_rcp")
        return out[0]
    cb = _RCP_FN(rcp)
    _rcp_callbacks.append(cb)          # keep the trampoline alive
    oracle.oracle_set_device_rcp_fn(cb)


def melt_like_positions(n, L, rng, jitter=0.12):
    """fcc lattice + Gaussian displacement, wrapped into (-L/2, L/2): a cheap stand-in for a melted fluid with
    no overlapping particles (used where the test does not integrate)."""
    L = np.asarray(L, dtype=np.float64)
    r = fcc_positions(n, L) + rng.normal(scale=jitter, size=(n, 3))
    r -= np.rint(r / L) * L
    return r


def pack_state(pos, species=None, vel=None, mass=None):
    """float4 arrays in HALMD's packing: position/velocity (2A,4) hi then lo planes; ghosts species -1."""
    n = pos.shape[0]
    A = array_size(n)
    p = np.zeros((2 * A, 4), dtype=np.float32)
    v = np.zeros((2 * A, 4), dtype=np.float32)
    hi = pos.astype(np.float32)
    lo = (pos - hi.astype(np.float64)).astype(np.float32)
    p[:n, :3] = hi
    p[A:A + n, :3] = lo
    sp = np.zeros(A, dtype=np.uint32)
    if species is not None:
        sp[:n] = species
    sp[n:] = PH
    p[:A, 3] = sp.view(np.float32)
    if vel is not None:
        vh = vel.astype(np.float32)
        v[:n, :3] = vh
        v[A:A + n, :3] = (vel - vh.astype(np.float64)).astype(np.float32)
    v[:A, 3] = 1.0
    if mass is not None:
        v[:n, 3] = mass
    return p, v, A


def species_of(p, A):
    return p[:A, 3].view(np.uint32)


def blocking_reference(n_steps, dt, max_lag, resolution, size, shift, separation, start=0, finalise=True):
    """Pure-Python restatement of blocking_scheme.cpp:75-200 (schedule only): returns the list of
    (level, [sample steps of the block]) correlation events.  Pinned against the reference's own classes compiled in
    place (tests/test_oracle_ref.py::test_blocking_schedule_against_reference_classes)."""
    max_interval = int(round(max_lag / dt))
    s = int(round(resolution / dt))
    interval = []
    while s <= max_interval:
        interval.append(s)
        if shift > 1 and s * shift <= max_interval:
            interval.append(s * shift)
        s *= size
    origin = [start] * len(interval)
    blocks = [[] for _ in interval]
    events = []

    def process(i):
        events.append((i, list(blocks[i])))
        skip = max((separation + interval[i] - 1) // interval[i], 1)
        origin[i] += skip * interval[i]
        del blocks[i][:skip]

    for step in range(start, start + n_steps + 1):
        for i, iv in enumerate(interval):
            if step >= origin[i] and (step - origin[i]) % iv == 0:
                blocks[i].append(step)
                if len(blocks[i]) == size:
                    process(i)
    for i in range(len(interval) if finalise else 0):
        while blocks[i]:
            process(i)
    return interval, events


def ref_blocking_schedule(ref, dt, max_lag, resolution, size, shift, separation, start, n_steps, finalise=True):
    """The reference's own clock + blocking_scheme + samples::blocking_scheme + correlation<> classes (oracle/ref_dynamics.cpp)
    driven like sampler:run; returns (intervals, events[(level, first, second)], count[level, lag], mean lag in steps, time grid)."""
    max_levels, max_events = 64, 4 * (n_steps + 2) * size + 1024
    iv = np.zeros(max_levels, dtype=np.uint64)
    ev = np.zeros((max_events, 3), dtype=np.uint64)
    nev = C.c_uint64(0)
    cnt, mean, time = (np.zeros(max_levels * size) for _ in range(3))
    fn = ref.ref_blocking_schedule
    fn.restype = C.c_int
    fn.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint, C.c_uint, C.c_uint, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p,
                   C.c_uint, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_void_p]
    levels = fn(dt, max_lag, resolution, size, shift, separation, start, n_steps, 1 if finalise else 0, iv.ctypes.data, max_levels,
                ev.ctypes.data, max_events, C.byref(nev), cnt.ctypes.data, mean.ctypes.data, time.ctypes.data)
    if levels < 0:
        return None
    n = levels * size
    return ([int(x) for x in iv[:levels]], ev[:nev.value].astype(np.int64), cnt[:n].reshape(levels, size),
            mean[:n].reshape(levels, size), time[:n].reshape(levels, size))


def blocking_events(events):
    """(level, [steps]) blocks of blocking_reference as (level, first, second) triples in correlation order."""
    return np.array([(lev, blk[0], st) for lev, blk in events for st in blk], dtype=np.int64).reshape(-1, 3)
