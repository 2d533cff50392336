"""Developer check: the chain-local step against the full-SpMM step on the same Philox streams.
usage: chain_local_check.py [small|mid|c3] [n_chains] [n_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

which = sys.argv[1] if len(sys.argv) > 1 else "small"
n_chains = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
if which == "small":
    ing = problems.tomo_small()
elif which == "mid":
    ing = synthetic.tomography_2d(nx=64, ny=48, n_sources_per_side=6, n_receivers=20, kmin=3, kmax=300)
else:
    ing = synthetic.tomography_2d()
spec = problems.spec_from_ingredients(ing)

os.environ["BBB_FULL_FORWARD"] = "1"
ref = Engine(spec, n_chains, seed=5)
os.environ["BBB_FULL_FORWARD"] = "0"
loc = Engine(spec, n_chains, seed=5)
print("chain_local flags:", ref.chain_local, loc.chain_local, flush=True)
ref.init_states()
loc.init_states()
torch.cuda.synchronize()
print("init ok", flush=True)
alive = np.ones(n_chains, bool)
for t in range(n_steps):
    ref.step(1)
    loc.step(1)
    torch.cuda.synchronize()
    mv0, mv1 = ref.move.cpu().numpy(), loc.move.cpu().numpy()
    a0, a1 = ref.accepted.cpu().numpy(), loc.accepted.cpu().numpy()
    l0, l1 = ref.log_like_ratio.cpu().numpy(), loc.log_like_ratio.cpu().numpy()
    assert np.array_equal(mv0[alive], mv1[alive]), t
    d = np.abs(l0 - l1)[alive]
    fin = np.isfinite(l0[alive]) & np.isfinite(l1[alive])
    worst = d[fin].max() if fin.any() else 0.0
    flip = alive & (a0 != a1)
    if flip.any():
        print("step", t, "decision flips", np.flatnonzero(flip), l0[flip], l1[flip])
        alive &= ~flip
    if t % 10 == 0 or worst > 1e-8:
        print("step %d: max |dllr| %.3e  alive %d" % (t, worst, alive.sum()), flush=True)
i0 = ref.current_cell_idx(0)
i1 = loc.current_cell_idx(0)
same = (i0 == i1).all(0).cpu().numpy()
print("index equal on alive chains:", same[alive].all(), " (alive %d / %d)" % (alive.sum(), n_chains))
m0, _ = ref.misfit_logdet()
m1, _ = loc.misfit_logdet()
rel = ((m0 - m1).abs() / m0.abs()).cpu().numpy()
print("misfit rel diff max (alive):", rel[alive].max())
loc.resync()
m2, _ = loc.misfit_logdet()
print("misfit drift vs resync:", ((m1 - m2).abs() / m2.abs()).max().item())
