#!/usr/bin/env python
"""Estimate for the "distance-shell rows" step of DESIGN.md §7 (CPU only, oracle host backend + scipy).

A neighbour-list entry (i, j) stored at build time with distance r_build can lie inside the cutoff r_c at a later step
only if r_build <= r_c + 2 d_max(t), d_max(t) being the largest displacement of any particle since the build (the
quantity the rebuild criterion already keeps on the device).  With each row ordered by shells of r_build, the force
kernel needs only the prefix of shells that can matter.  This script runs a melted LJ liquid (config-2 state point) on
the host backend and reports, averaged over the steps of complete rebuild intervals, which fraction of the stored
entries would be walked — exactly (sorted rows) and for a few shell counts — and verifies on every step that no pair
inside r_c is ever excluded."""
import ctypes as C
import json
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import support  # noqa: E402

oracle = support.load_oracle()
oracle.host_sim_create.restype = C.c_void_p
oracle.host_sim_rebuilds.restype = C.c_size_t
n, rho, rc, skin, dt = 4000, 0.8442, 2.5, 0.5, 0.001
L = (n / rho) ** (1 / 3)
rng = np.random.default_rng(1)
pos = np.ascontiguousarray(support.fcc_positions(n, [L] * 3))
vel = rng.normal(scale=np.sqrt(1.2 * 1.9), size=(n, 3))        # melts to T ~ 1.2
vel = np.ascontiguousarray(vel - vel.mean(axis=0))
sim = C.c_void_p(oracle.host_sim_create(C.c_size_t(n), support.d3([L] * 3), 1, support.darr([1.0]), support.darr([1.0]),
                                        support.darr([rc]), 2, C.c_double(0.005), C.c_double(skin), C.c_double(dt), 8))
oracle.host_sim_set(sim, support.ptr(pos), support.ptr(vel), None, None)
oracle.host_sim_run(sim, C.c_size_t(3000))


def state():
    p, v, f = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    img = np.zeros((n, 3), dtype=np.int32)
    oracle.host_sim_get(sim, support.ptr(p), support.ptr(v), support.ptr(f), support.ptr(img))
    return p


def wrap(d):
    return d - np.rint(d / L) * L


shell_counts = (2, 4, 8)
acc = {"exact": [], **{k: [] for k in shell_counts}}
intervals, violations = [], 0
rebuilds = oracle.host_sim_rebuilds(sim)
r0 = pairs = r_build = None
steps_in = 0
for step in range(4000):
    oracle.host_sim_run(sim, C.c_size_t(1))
    p = state()
    rb = oracle.host_sim_rebuilds(sim)
    if rb != rebuilds:                                  # lists were rebuilt for these positions
        rebuilds = rb
        if r0 is not None:
            intervals.append(steps_in)
        r0 = p.copy()
        tree = cKDTree(np.mod(r0, L), boxsize=L)
        pairs = tree.query_pairs(rc + skin, output_type="ndarray")
        r_build = np.linalg.norm(wrap(r0[pairs[:, 0]] - r0[pairs[:, 1]]), axis=1)
        steps_in = 0
        continue
    if r0 is None:
        continue
    steps_in += 1
    d_max = np.linalg.norm(wrap(p - r0), axis=1).max()
    reach = rc + 2 * d_max
    r_now = np.linalg.norm(wrap(p[pairs[:, 0]] - p[pairs[:, 1]]), axis=1)
    violations += int(np.count_nonzero((r_now < rc) & (r_build > reach)))
    acc["exact"].append(np.count_nonzero(r_build <= reach) / len(r_build))
    for k in shell_counts:
        edges = rc + skin * np.arange(1, k + 1) / k                      # upper bounds of the shells
        upto = edges[np.searchsorted(edges, min(reach, rc + skin))]      # whole shells only
        acc[k].append(np.count_nonzero(r_build <= upto) / len(r_build))
out = {"n": n, "steps_per_rebuild": float(np.mean(intervals)), "pairs_inside_cutoff_excluded": violations,
       "fraction_within_cutoff": float(np.mean(r_now < rc)),
       "entries_walked_exact": float(np.mean(acc["exact"])),
       **{f"entries_walked_{k}_shells": float(np.mean(acc[k])) for k in shell_counts}}
print(json.dumps(out))
