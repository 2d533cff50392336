"""Developer timing probe (not the bench): C3-shape chain-local engine, (a) one iteration per call with the L2 flushed
between calls (the bench's `value` protocol), (b) the same iterations back to back in one call.
usage: step_time.py [chains] [burn] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 300
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
ing = synthetic.tomography_2d()
spec = problems.spec_from_ingredients(ing)
eng = Engine(spec, n_chains, seed=12345)
eng.init_states()
eng.step(burn)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for rep in range(3):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        eng.step(1)
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.step(steps)
    e1.record()
    torch.cuda.synchronize()
    out.append((ms.mean(), e0.elapsed_time(e1) / steps))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.step(1024)
e1.record()
torch.cuda.synchronize()
print("1024 iterations in one call: %.4f ms/step" % (e0.elapsed_time(e1) / 1024))
fl = min(o[0] for o in out)
bb = min(o[1] for o in out)
print("flushed %.4f ms/step (%.2f M/s)   back-to-back %.4f ms/step (%.2f M/s)   all: %s" % (
    fl, n_chains / fl / 1e3, bb, n_chains / bb / 1e3, " ".join("%.4f/%.4f" % o for o in out)), flush=True)
