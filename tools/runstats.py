"""Where the blocks of the fused persistent kernel (k_chain_run) spend their time (developer probe; needs
`make -C bayesbay_b200/csrc timeline` and BBB_LIB_PATH=bayesbay_b200/libbbb200_tl.so).
usage: runstats.py [chains] [burn] [iterations per call] [c5]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bayesbay_b200 import _lib, synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 300
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ing = synthetic.tomography_2d()
spec = problems.spec_from_ingredients(ing)
eng = Engine(spec, n_chains, seed=12345)
eng.init_states()
eng.step(burn)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((1024, 16), np.int64)
lib.bbb_debug_runstats(buf.ctypes.data_as(C.c_void_p), 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.step(iters)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
lib.bbb_debug_runstats(buf.ctypes.data_as(C.c_void_p), 1)
b = buf[buf[:, 12] > 0].astype(float)
mhz = 1965.0
print("%d iterations of %d chains in %.3f ms (%.4f ms per iteration); %d blocks" % (iters, n_chains, ms, ms / iters, len(b)))
tot = b[:, 12].mean() / mhz
print("worker loop per block: %.1f us; items per block %.1f worked + %.1f decided by the proposer" % (tot, b[:, 11].mean(), b[:, 5].mean()))
print("workers:  wait for a record %.1f us (%.1f %%), body %.1f us (%.2f us per item), epilogue %.1f us (%.2f per item)" % (
    b[:, 8].mean() / mhz, 100 * b[:, 8].sum() / b[:, 12].sum(), b[:, 9].mean() / mhz, b[:, 9].sum() / b[:, 11].sum() / mhz,
    b[:, 10].mean() / mhz, b[:, 10].sum() / b[:, 11].sum() / mhz))
print("proposer: wait for the chain's previous iteration %.1f us, proposals %.1f us (%.2f us each), records %.1f us (%.2f each), "
      "own decisions %.1f us (%.2f each), wait for a free record %.1f us" % (
          b[:, 0].mean() / mhz, b[:, 1].mean() / mhz, b[:, 1].sum() / b[:, 4].sum() / mhz, b[:, 2].mean() / mhz,
          b[:, 2].sum() / b[:, 11].sum() / mhz, b[:, 6].mean() / mhz, b[:, 6].sum() / max(b[:, 5].sum(), 1) / mhz, b[:, 3].mean() / mhz))
