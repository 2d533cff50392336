"""Short C3-shaped run for ncu captures: init, burn a little, then a few staged steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 200
spec = problems.spec_from_ingredients(synthetic.tomography_2d())
eng = Engine(spec, n_chains, seed=1)
eng.init_states()
eng.step(burn)
torch.cuda.synchronize()
for _ in range(6):
    for s in range(4):
        eng.stage(s)
torch.cuda.synchronize()
print("ok", eng.gather_space(0)[0].float().mean().item())
