import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from halmd_b200 import mdsim, observables, simulation
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
s = simulation.kob_andersen_mixture(n, seed=1)
s.integrator.set_timestep(0.005); s.run(300); s.integrator.set_timestep(0.001)
def T(f, label):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); print(f"{label}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
T(lambda: s.run(500), "run 500 (no sampling)")
T(lambda: s.run(500), "run 500 (no sampling)")
L = max(s.box.length)
qgrid = observables.utility.semilog_grid(start=2*np.pi/L, stop=15.0, decimation=10).value
T(lambda: None, "nop")
t0=time.perf_counter(); wv = observables.utility.wavevector(box=s.box, wavenumber=qgrid, tolerance=0.01, max_count=7); print("wavevector picker", time.perf_counter()-t0, len(wv.value))
bs = observables.dynamics.blocking_scheme(max_lag=1.4, every=100, size=10, wavevector=wv)
for k in range(3):
    T(lambda: bs.sample(), "bs.sample()")
    T(lambda: s.run(100), "run 100")
T(lambda: observables.sampler.run(500, blocking_scheme=bs), "sampler.run 500 with blocking")
T(lambda: observables.sampler.run(1000, blocking_scheme=bs), "sampler.run 1000 with blocking")
bs2 = observables.dynamics.blocking_scheme(max_lag=1.4, every=100, size=10)
T(lambda: observables.sampler.run(1000, blocking_scheme=bs2), "sampler.run 1000 with blocking, no wavevectors")
