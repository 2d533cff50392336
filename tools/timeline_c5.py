"""Coarse phase timeline of k_chain_step on the C5 shape (multi-tile path): developer probe, needs the timeline build
(`make -C bayesbay_b200/csrc timeline`, BBB_LIB_PATH=bayesbay_b200/libbbb200_tl.so)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesbay_b200 import synthetic, _lib
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 592
ing = synthetic.tomography_2d(nx=256, ny=256, n_sources_per_side=100, n_receivers=250, kmin=50, kmax=500)
spec = problems.spec_from_ingredients(ing)
eng = Engine(spec, n_chains, seed=1)
eng.init_states()
eng.step(100)
torch.cuda.synchronize()
lib = _lib.load()
acc = []
for rep in range(6):
    eng.stage(0)
    lib.bbb_debug_timeline_clear()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.stage(1); eng.stage(2); eng.stage(3); b.record()
    torch.cuda.synchronize()
    tl = np.zeros((n_chains, 24), np.int64)
    lib.bbb_debug_timeline(tl.ctypes.data_as(C.c_void_p), n_chains)
    acc.append(tl)
    print("launch %.1f us" % (a.elapsed_time(b) * 1e3))
t = np.concatenate(acc[2:])
t = t[t[:, 7] > 0]
mhz = 1965.0
seg = [("entry->tables", 0, 1), ("tables->raster", 1, 2), ("raster->rederive", 2, 3), ("tile loop (all tiles: prep, columns, park)", 3, 11),
       ("block sum", 11, 6), ("decision", 6, 7)]
for nm, i, j in seg:
    d = (t[:, j] - t[:, i]) / mhz
    print("%-45s mean %7.2f us  p50 %7.2f  p90 %7.2f" % (nm, d.mean(), np.median(d), np.percentile(d, 90)))
accd = t[:, 14] > 0
print("commit (accepted only, %d) mean %.2f us" % (accd.sum(), ((t[accd, 8] - t[accd, 7]) / mhz).mean()))
print("last tile: prep->columns done %.2f us" % (((t[:, 5] - t[:, 4]) / mhz).mean()))
wall = (t[:, 17] - t[:, 16]) / 1e3
print("block wall mean %.1f us p90 %.1f max %.1f; n_list mean %.0f p90 %.0f; queued groups mean %.1f; k mean %.0f" % (
    wall.mean(), np.percentile(wall, 90), wall.max(), t[:, 12].mean(), np.percentile(t[:, 12], 90), t[:, 13].mean(), t[:, 15].mean()))
