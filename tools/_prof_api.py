import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayesbay_b200 as bb
from bayesbay_b200 import synthetic
from bayesbay_b200.forward import RayTomographyForward
ing = synthetic.tomography_2d()
fwd = RayTomographyForward((ing["indptr"], ing["indices"], ing["data"]), ing["points"], "voronoi", "vel")
par, ll = synthetic.build_problem(ing, bb, lambda _: fwd)
inv = bb.BayesianInversion(par, ll, n_chains=1024)
pr = cProfile.Profile(); pr.enable()
inv.run(n_iterations=2000, burnin_iterations=1000, save_every=100, verbose=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
