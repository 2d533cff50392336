"""Developer probe: time of one parallel-tempering swap round (bbb_pt_swap) for n chains."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesbay_b200._engine import pt_swap
for n in (1024, 2048, 8192, 16384):
    rng = np.random.default_rng(0)
    g = torch.as_tensor(np.stack((rng.uniform(50, 150, n), rng.uniform(-20, 20, n), np.tile(np.geomspace(1, 5, 16), n // 16)), axis=1).copy()).cuda()
    for _ in range(3):
        pt_swap(g.clone(), 7, 1, 0, n)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gg = [g.clone() for _ in range(10)]
    a.record()
    for x in gg:
        pt_swap(x, 7, 1, 0, n)
    b.record(); torch.cuda.synchronize()
    print(n, "chains: %.3f ms per swap round" % (a.elapsed_time(b) / 10))
