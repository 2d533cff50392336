"""Developer probe: ms/step of multi-iteration calls of different lengths (C3 shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

spec = problems.spec_from_ingredients(synthetic.tomography_2d())
eng = Engine(spec, 1024, seed=12345)
eng.init_states()
eng.step(300)
torch.cuda.synchronize()
def t(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for n in (16, 50, 100, 200, 255, 256, 512, 1024, 50, 50):
    since = int(eng.lib.bbb_engine_steps_since_resync(eng._handle))
    k = eng.gather_space(0)[0].float().mean().item()
    print("n %5d  since_resync %4d  k_mean %.1f  ->  %.4f ms/step" % (n, since, k, t(n)), flush=True)
import ctypes as C
eng.lib.bbb_engine_set_resync(eng._handle, 0, 0)
for n in (256, 1024):
    print("no resync: n %5d -> %.4f ms/step" % (n, t(n)), flush=True)
