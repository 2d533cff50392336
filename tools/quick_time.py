"""Developer timing probe (not the bench): C3-shape engine step rate with CUDA events."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
t0 = time.time()
ing = synthetic.tomography_2d()
spec = problems.spec_from_ingredients(ing)
print("problem built in %.1fs nnz=%d" % (time.time() - t0, ing["indices"].size), flush=True)
eng = Engine(spec, n_chains, seed=1)
eng.init_states()
torch.cuda.synchronize()
eng.step(10)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("rep %d: %.3f ms/step  %.0f chain-steps/s" % (rep, ms / steps, n_chains * steps / ms * 1e3), flush=True)
k = eng.gather_space(0)[0].float()
print("k mean %.1f  acc rate %.3f" % (k.mean().item(), eng.n_accepted.sum().item() / eng.n_proposed.sum().item()))
p = eng.n_proposed.sum(1).cpu().numpy(); a = eng.n_accepted.sum(1).cpu().numpy()
print("proposed", p, "acc", np.round(a / np.maximum(p, 1), 3))
