"""Short C3-shaped run for ncu captures of the fused persistent kernel: init, burn, then a few multi-iteration calls.
usage: profile_run.py [chains] [burn] [iterations per call]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 200
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 16
spec = problems.spec_from_ingredients(synthetic.tomography_2d())
eng = Engine(spec, n_chains, seed=1)
eng.init_states()
eng.step(burn)
torch.cuda.synchronize()
for _ in range(4):
    eng.step(iters)
torch.cuda.synchronize()
print("ok", eng.gather_space(0)[0].float().mean().item())
