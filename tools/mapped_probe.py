"""Developer probe of the mapped hosted step: host-side and device-side time per call.  usage: mapped_probe.py chains burn steps"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 300
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
spec = problems.spec_from_ingredients(synthetic.tomography_2d())
eng = Engine(spec, n_chains, seed=1)
eng.init_states()
eng.step(burn)
rw = eng.hosted_record_words()
host = [torch.zeros((eng.C, rw), dtype=torch.int64).pin_memory() for _ in range(2)]
eng.pack_states_records(host[0])
torch.cuda.synchronize()
cur = 0
for _ in range(5):
    eng.step_hosted_mapped(host[cur], host[cur ^ 1]); eng.step_hosted_mapped_wait(); cur ^= 1
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
t_enq = t_wait = t_gap = 0.0
t_prev = time.perf_counter()
for i in range(steps):
    t0 = time.perf_counter()
    ev[i][0].record()
    eng.step_hosted_mapped(host[cur], host[cur ^ 1])
    ev[i][1].record()
    t1 = time.perf_counter()
    eng.step_hosted_mapped_wait()
    t2 = time.perf_counter()
    cur ^= 1
    t_enq += t1 - t0; t_wait += t2 - t1; t_gap += t0 - t_prev; t_prev = t2
torch.cuda.synchronize()
dev = sum(a.elapsed_time(b) for a, b in ev) / steps * 1e3
print("per call: host enqueue %.1f us, host wait %.1f us, host between calls %.1f us; device (chain + next proposal, event to event) %.1f us"
      % (t_enq / steps * 1e6, t_wait / steps * 1e6, t_gap / steps * 1e6, dev))
# resident comparison: chain kernel alone via stage()
torch.cuda.synchronize()
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
tp = tc = 0.0
for i in range(50):
    a.record(); eng.stage(0); b.record(); eng.stage(2); c.record(); torch.cuda.synchronize()
    tp += a.elapsed_time(b); tc += b.elapsed_time(c)
print("resident staged: propose %.1f us, chain %.1f us" % (tp / 50 * 1e3, tc / 50 * 1e3))
