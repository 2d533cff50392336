"""C5-shape probe (BASELINE configs[4]): 256x256 grid, 100k rays, K=500, 1024 chains."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t0 = time.time()
ing = synthetic.tomography_2d(nx=256, ny=256, n_sources_per_side=100, n_receivers=250, kmin=50, kmax=500)
print("C5 problem: rays %d nnz %d (%.1f/row) built in %.0fs" % (ing["indptr"].size - 1, ing["indices"].size,
      ing["indices"].size / (ing["indptr"].size - 1), time.time() - t0), flush=True)
spec = problems.spec_from_ingredients(ing)
eng = Engine(spec, n_chains, seed=1)
eng.init_states()
torch.cuda.synchronize()
print("mem GB", torch.cuda.memory_allocated() / 1e9, flush=True)
eng.step(50)
torch.cuda.synchronize()
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(20); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("C5: %.3f ms/step  %.0f chain-steps/s" % (ms, n_chains / ms * 1e3), flush=True)
names = ["propose", "raster", "spmm", "finish"]; tot = np.zeros(4)
for i in range(10):
    for s in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.stage(s); b.record(); torch.cuda.synchronize(); tot[s] += a.elapsed_time(b)
print(dict(zip(names, (tot / 10).round(3))))
m0, _ = eng.misfit_logdet(); eng.refresh(); m1, _ = eng.misfit_logdet()
print("cached vs fresh misfit max rel err", ((m0 - m1).abs() / m1.abs()).max().item(), "k mean", eng.gather_space(0)[0].float().mean().item())
