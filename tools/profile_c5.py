"""Short C5-shaped run for ncu captures (BASELINE configs[4]; 256 chains keep the capture short)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ing = synthetic.tomography_2d(nx=256, ny=256, n_sources_per_side=100, n_receivers=250, kmin=50, kmax=500)
eng = Engine(problems.spec_from_ingredients(ing), n_chains, seed=1)
eng.init_states()
eng.step(150)
torch.cuda.synchronize()
for _ in range(4):
    for s in range(4):
        eng.stage(s)
torch.cuda.synchronize()
print("ok", eng.gather_space(0)[0].float().mean().item())
