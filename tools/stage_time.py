"""Developer timing probe: CUDA-event time of each stage of the C3 step (after a burn-in)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 300
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
spec = problems.spec_from_ingredients(synthetic.tomography_2d())
eng = Engine(spec, n_chains, seed=1)
eng.init_states()
eng.step(burn)
torch.cuda.synchronize()
tot = [0.0] * 4
for _ in range(reps):
    for s in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.stage(s); b.record(); torch.cuda.synchronize()
        tot[s] += a.elapsed_time(b)
print("chain_local", eng.chain_local, "stage ms:", [round(t / reps, 4) for t in tot], "sum", round(sum(tot) / reps, 4))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.step(200); b.record(); torch.cuda.synchronize()
print("back-to-back ms/step", a.elapsed_time(b) / 200)
