"""Developer probe: wall time of the public-API run (BayesianInversion.run + get_results) on the C3 problem, after a
warm-up run (lazy module loading, graph capture, pinned buffers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayesbay_b200 as bb
from bayesbay_b200 import synthetic

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
save_every = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ing = synthetic.tomography_2d()
par, ll = synthetic.build_problem(ing, bb)
inv = bb.BayesianInversion(par, ll, n_chains=n_chains)
inv.run(n_iterations=600, burnin_iterations=100, save_every=save_every, verbose=False)
inv.get_results()
torch.cuda.synchronize()
for rep in range(2):
    t1 = time.perf_counter()
    inv.run(n_iterations=n_it, burnin_iterations=0, save_every=save_every, verbose=False)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res = inv.get_results()
    t3 = time.perf_counter()
    print("run %.3fs  get_results %.3fs  -> %.2f M chain-steps/s incl. saves and results; samples %d" % (
        t2 - t1, t3 - t2, n_chains * n_it / (t3 - t1) / 1e6, len(res[list(res)[0]])), flush=True)
# where the time goes: stepping alone
t1 = time.perf_counter(); inv.batch.engine.step(n_it); torch.cuda.synchronize(); t2 = time.perf_counter()
print("stepping alone %.3fs" % (t2 - t1))
