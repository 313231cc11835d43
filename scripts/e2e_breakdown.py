#!/usr/bin/env python
"""Development aid: where the end-to-end (host buffers) leg of bench.py spends its time.  Run on a GPU box."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halmd_b200 import simulation  # noqa: E402

s = simulation.lennard_jones_fluid(256000, seed=1, integrator="andersen", timestep=0.005, rate=2.0)
s.run(600)
from halmd_b200 import mdsim  # noqa: E402
mdsim.integrators.verlet(box=s.box, particle=s.particle, timestep=0.001)
s.core = mdsim.core.get()
s.run(1000)
pos = s.particle.get("position_ds")
vel = s.particle.get("velocity_ds")


def timed(label, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
    return out


acc = {}
reps = 4
for _ in range(reps):
    timed("set position_ds", lambda: s.particle.set("position_ds", pos), acc)
    timed("set velocity_ds", lambda: s.particle.set("velocity_ds", vel), acc)
    timed("run 1000", lambda: s.run(1000), acc)
    pos = timed("get position_ds", lambda: s.particle.get("position_ds"), acc)
    vel = timed("get velocity_ds", lambda: s.particle.get("velocity_ds"), acc)
    timed("thermodynamics", lambda: s.thermodynamics(), acc)
for k, v in acc.items():
    print(f"{k:18s} {v / reps:8.2f} ms")
