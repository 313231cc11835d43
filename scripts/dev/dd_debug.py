"""Debug aid: decomposed system on shared GPU(s) with gloo ranks; prints the x-extent and cell layers per rank."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from halmd_b200 import simulation
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank % torch.cuda.device_count())
backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
dist.init_process_group(backend, device_id=torch.device("cuda", rank % torch.cuda.device_count()) if backend == "nccl" else None)
n = int(sys.argv[1])
dd = simulation.decomposed_lj_fluid(n, rank, world, seed=1, integrator="andersen", timestep=0.005, rate=2.0)
try:
    dd.run(int(sys.argv[2]))
    print(rank, "ok", dd.updates, flush=True)
except Exception as e:
    print(rank, "ERR", str(e)[:200], flush=True)
    torch.cuda.synchronize()
    ntot = dd.n_own + dd.n_ghost
    x = dd.pos[:ntot, 0].float().cpu().numpy()
    w = dd.pos[:ntot, 3].view(torch.int32).cpu().numpy()
    ell, nx = float(dd.cell_length[0]), int(dd.ncell[0])
    c = np.floor(x.astype(np.float64) / ell + nx).astype(np.int64) % nx
    rel = (c - dd.cell_x0) % nx
    bad = np.nonzero(rel >= dd.cell_nx)[0]
    print(rank, "n_own", dd.n_own, "n_ghost", dd.n_ghost, "lo/hi", float(dd.dec.lo), float(dd.dec.hi), "layers", dd.cell_x0, dd.cell_nx,
          "bad", len(bad), bad[:5], x[bad[:5]], w[bad[:5]], "owned x", x[:dd.n_own].min(), x[:dd.n_own].max(), flush=True)

