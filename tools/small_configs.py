"""Throughput of the thread-per-chain fused kernel on BASELINE configs[0] and [1] (documentation only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

for name, ing, chains, steps in (("C1 partition_1d, 4 chains", synthetic.partition_1d(), 4, 20000),
                                 ("C1 partition_1d, 4096 chains", synthetic.partition_1d(), 4096, 5000),
                                 ("C2 polyfit, 4096 chains", synthetic.polyfit(), 4096, 20000),
                                 ("C2 polyfit, 65536 chains", synthetic.polyfit(), 65536, 5000)):
    spec = problems.spec_from_ingredients(ing)
    eng = Engine(spec, chains, seed=3)
    eng.init_states()
    eng.step(200)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    acc = eng.n_accepted.sum().item() / eng.n_proposed.sum().item()
    print(f"{name}: {chains * steps / ms * 1e3:,.0f} chain-steps/s ({ms / steps * 1e3:.2f} us per step of all chains), acceptance {acc:.3f}")
