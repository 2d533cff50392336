"""Developer probe: wall time of the public-API run (BayesianInversion.run) on the C3 problem."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayesbay_b200 as bb
from bayesbay_b200 import synthetic
from bayesbay_b200.forward import RayTomographyForward

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
save_every = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ing = synthetic.tomography_2d()
fwd = RayTomographyForward((ing["indptr"], ing["indices"], ing["data"]), ing["points"], "voronoi", "vel")
par, ll = synthetic.build_problem(ing, bb, lambda _: fwd)
t0 = time.perf_counter()
inv = bb.BayesianInversion(par, ll, n_chains=n_chains)
t1 = time.perf_counter()
inv.run(n_iterations=n_it, burnin_iterations=n_it // 2, save_every=save_every, verbose=False)
torch.cuda.synchronize()
t2 = time.perf_counter()
res = inv.get_results()
t3 = time.perf_counter()
print("construct %.2fs  run %.2fs (%.0f chain-steps/s incl. saves)  get_results %.2fs  samples %d" % (
    t1 - t0, t2 - t1, n_chains * n_it / (t2 - t1), t3 - t2, len(res[list(res)[0]])))
