"""Developer probe: the UNMODIFIED reference's BayesianInversion.run driving the CUDA engine through the plug-in samplers
(bayesbay_b200/plugin.py), host wall clock by phase.  usage: plugin_time.py n_chains n_iterations save_every"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesbay_b200 import synthetic
from bayesbay_b200.plugin import cuda_samplers
from oracle import reference_runner as RR

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
save_every = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ing = synthetic.tomography_2d()
bayesbay = RR.import_reference()
t0 = time.perf_counter()
par, ll = RR.build_problem(bayesbay, ing)
inv = bayesbay.BayesianInversion(par, ll, n_chains=n_chains)
t1 = time.perf_counter()
print("reference builds and initialises %d chains: %.2f s" % (n_chains, t1 - t0))
CudaVanillaSampler, _ = cuda_samplers(bayesbay, seed=3)
for rep in range(2):
    t1 = time.perf_counter()
    inv.run(sampler=CudaVanillaSampler(), n_iterations=n_it, burnin_iterations=0, save_every=save_every, verbose=False)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res = inv.get_results()
    t3 = time.perf_counter()
    print("run %d: inversion.run %.3f s (%.2f M chain-steps/s), get_results %.3f s, %d samples"
          % (rep, t2 - t1, n_chains * n_it / (t2 - t1) / 1e6, t3 - t2, len(res["voronoi.n_dimensions"])))
