"""Phase timeline of k_chain_step (developer probe; needs `make -C bayesbay_b200/csrc timeline` and
BBB_LIB_PATH=bayesbay_b200/libbbb200_tl.so): SM-clock stamps at the phase boundaries of every block of one launch."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesbay_b200 import synthetic, _lib
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 300
spec = problems.spec_from_ingredients(synthetic.tomography_2d())
eng = Engine(spec, n_chains, seed=1)
eng.init_states()
eng.step(burn)
torch.cuda.synchronize()
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = ["entry->tables", "tables->raster", "raster->rederive", "rederive->prep", "prep->accum", "accum->misfit(sum)",
         "sum->decided", "decided->commit end"]
acc = []
for rep in range(8):
    flush.fill_(rep)
    eng.stage(0)
    lib.bbb_debug_timeline_clear()
    torch.cuda.synchronize()
    eng.stage(1); eng.stage(2); eng.stage(3)
    torch.cuda.synchronize()
    tl = np.zeros((n_chains, 24), np.int64)
    lib.bbb_debug_timeline(tl.ctypes.data_as(C.c_void_p), n_chains)
    acc.append(tl)
tl = np.concatenate(acc[2:])
work = tl[:, 7] > 0
t = tl[work]
print("blocks with a forward: %d of %d" % (work.sum(), len(tl)))
mhz = 1965.0
d = np.diff(t[:, :8], axis=1) / mhz
for i, nm in enumerate(names[:7]):
    print("%-22s mean %6.2f us  p50 %6.2f  p90 %6.2f" % (nm, d[:, i].mean(), np.median(d[:, i]), np.percentile(d[:, i], 90)))
accd = t[:, 14] > 0
print("%-22s mean %6.2f us (accepted only, %d)" % (names[7], ((t[accd, 8] - t[accd, 7]) / mhz).mean(), accd.sum()))
tot = np.where(accd, t[:, 8], t[:, 7]) - t[:, 0]
print("block total mean %.2f us p50 %.2f p90 %.2f max %.2f" % (tot.mean() / mhz, np.median(tot) / mhz, np.percentile(tot, 90) / mhz, tot.max() / mhz))
print("warp-15 decision prep ends %.2f us after entry (raster barrier at %.2f; thread 0 reaches it at %.2f)" % (
    ((t[:, 9] - t[:, 0]) / mhz).mean(), ((t[:, 2] - t[:, 0]) / mhz).mean(), ((t[:, 10] - t[:, 0]) / mhz).mean()))
print("thread 0 reaches the block sum %.2f us before it completes" % (((t[:, 6] - t[:, 11]) / mhz).mean()))
print("n_list mean %.1f  p90 %.0f max %d; queued groups mean %.1f; k_new mean %.1f; accepted %.3f" % (
    t[:, 12].mean(), np.percentile(t[:, 12], 90), t[:, 12].max(), t[:, 13].mean(), t[:, 15].mean(), accd.mean()))
print("sum of block totals / 296 slots = %.1f us per launch" % (tot.sum() / mhz / 296 / (len(acc) - 2)))

# scheduling: wall-clock (globaltimer) span of the launch against the busy time of the block slots
for tl1 in acc[2:5]:
    st, en = tl1[:, 16], tl1[:, 17]
    span = (en.max() - st.min()) / 1e3
    busy = (en - st).sum() / 1e3
    w = tl1[:, 7] > 0
    order = np.argsort(st)
    last_start = (st.max() - st.min()) / 1e3
    print("launch span %.1f us; sum of block wall times %.1f us (/296 = %.1f); forward blocks %d (mean wall %.1f us), others mean %.2f us; last block starts at %.1f us"
          % (span, busy, busy / 296, w.sum(), ((en - st)[w]).mean() / 1e3, ((en - st)[~w]).mean() / 1e3, last_start))
    # running blocks over time
    ts = np.linspace(st.min(), en.max(), 15)
    print("  running blocks over the launch:", [int(((st <= x) & (en > x)).sum()) for x in ts])
    sm = tl1[:, 18]
    print("  SMs used %d; blocks per SM min %d max %d" % (len(np.unique(sm)), np.bincount(sm).min(), np.bincount(sm).max()))

# block time by move and cell count (input of the longest-first block order)
tt = np.concatenate(acc[2:])
tt = tt[tt[:, 7] > 0]
dur = (tt[:, 17] - tt[:, 16]) / 1e3
inner = tt[:, 19] & 3
kk = tt[:, 15]
for mv, nm in enumerate(["birth", "death", "values", "sites"]):
    m = inner == mv
    if m.sum() < 10: continue
    x = 10000.0 / kk[m]
    A = np.stack([np.ones(m.sum()), x], 1)
    coef = np.linalg.lstsq(A, dur[m], rcond=None)[0]
    resid = dur[m] - A @ coef
    print("%-7s n %5d mean %.1f us  fit %.1f + %.4f * G/k  (rms resid %.1f)  n_list mean %.0f" % (nm, m.sum(), dur[m].mean(), coef[0], coef[1], resid.std(), tt[m, 12].mean()))
    for lo, hi in ((0, 20), (20, 35), (35, 50), (50, 70), (70, 100), (100, 201)):
        mm = m & (kk >= lo) & (kk < hi)
        if mm.sum() > 5: print("     k in [%3d,%3d): n %4d mean %.1f p90 %.1f" % (lo, hi, mm.sum(), dur[mm].mean(), np.percentile(dur[mm], 90)))

# phase durations by move (us, medians)
print("phase medians by move: entry, raster, rederive, prep, accum, misfit, decide | decision-prep end, n_list")
for mv, nm in enumerate(["birth", "death", "values", "sites"]):
    m = inner == mv
    if m.sum() < 10: continue
    dd = np.diff(tt[m][:, :8], axis=1) / mhz
    print("%-7s" % nm, " ".join("%6.2f" % np.median(dd[:, i]) for i in range(7)), "| %6.2f %5.0f" % (np.median((tt[m, 9] - tt[m, 0]) / mhz), np.median(tt[m, 12])))
