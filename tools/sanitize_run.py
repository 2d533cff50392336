"""Small chain-local run for compute-sanitizer (memcheck / racecheck / synccheck): resident steps (single and
multi-iteration calls), a from-scratch re-evaluation, the three hosted steps, a swap round and a save point.
usage: compute-sanitizer --tool <tool> python tools/sanitize_run.py [n_chains] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine, pt_swap
from tests import problems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ing = synthetic.tomography_2d(nx=24, ny=20, n_sources_per_side=4, n_receivers=8, kmin=4, kmax=40)
spec = problems.spec_from_ingredients(ing)
eng = Engine(spec, n, seed=11)
eng.init_states()
for _ in range(steps):
    eng.step(1)
eng.step(steps)
eng.resync()
K = spec["spaces"][0]["kmax"]
# mapped hosted step
rw = eng.hosted_record_words()
host = [torch.zeros((n, rw), dtype=torch.int64).pin_memory() for _ in range(2)]
eng.pack_states_records(host[0])
torch.cuda.synchronize()
for it in range(steps):
    eng.step_hosted_mapped(host[it & 1], host[(it + 1) & 1])
    assert eng.step_hosted_mapped_wait() == 0
# ragged and padded hosted steps (copy engine)
cap = eng.hosted_ragged_words()
hr = [torch.zeros((cap,), dtype=torch.int64, pin_memory=True) for _ in range(2)]
stage = torch.zeros((cap,), dtype=torch.int64, device="cuda"); out = torch.zeros((cap,), dtype=torch.int64, device="cuda")
hr[0].copy_(eng.pack_states_ragged()); torch.cuda.synchronize()
for it in range(3):
    eng.step_hosted_ragged(hr[it & 1], hr[(it + 1) & 1], stage, out)
    assert eng.step_hosted_wait() == 0
# swap round and results
rows = eng.misfit_logdet_T()
pt_swap(rows, 5, 0, 0, n, out=eng.temperature)
eng.step(2)
k, sites, vals = eng.gather_space(0)
mis, ld = eng.misfit_logdet()
torch.cuda.synchronize()
print("sanitize_run ok", int(k.max().item()), float(mis.mean().item()))
