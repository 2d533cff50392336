"""Developer timing probe: the bench's e2e leg (hosted step) alone, after a burn-in.  usage: e2e_time.py chains burn steps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine
from tests import problems

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 300
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
spec = problems.spec_from_ingredients(synthetic.tomography_2d())
eng = Engine(spec, n_chains, seed=1)
eng.init_states()
eng.step(burn)
torch.cuda.synchronize()
r = bench.e2e_run(eng, steps, torch.device("cuda:0"), 1, None)
print("pipeline", os.environ.get("BBB_PIPELINE", "4"), "pdl", os.environ.get("BBB_DEPENDENT_LAUNCH", "1"),
      "e2e %.3f M chain-steps/s  %.4f ms/step (host wall %.4f)  h2d %d d2h %d refreshes %d" % (
          r["value"] / 1e6, r["ms_per_step"], r["host_wall_ms_per_step"], r["h2d_bytes_per_step"], r["d2h_bytes_per_step"], r["cache_refreshes"]))
if os.environ.get("E2E_DETAIL"):
    import time
    K = eng.spec["spaces"][0]["kmax"]
    cap = eng.state_image_bytes(K) // 8
    host = [torch.empty((cap,), dtype=torch.int64, pin_memory=True) for _ in range(2)]
    stage = torch.empty((cap,), dtype=torch.int64, device="cuda"); out = torch.empty((cap,), dtype=torch.int64, device="cuda")
    kb = min(K, int(eng.gather_space(0)[0].max().item()))
    n0 = eng.state_image_bytes(kb) // 8
    host[0][:n0].copy_(eng.pack_states_hosted(kb)); torch.cuda.synchronize()
    cur = 0; enq = 0.0; tot = 0.0
    for it in range(steps + 3):
        t0 = time.perf_counter()
        kb = eng.step_hosted(kb, host[cur], host[cur ^ 1], stage, out)[0]
        t1 = time.perf_counter()
        eng.step_hosted_wait()
        t2 = time.perf_counter()
        cur ^= 1
        if it >= 3: enq += t1 - t0; tot += t2 - t0
    print("enqueue %.1f us  total %.1f us per call" % (enq / steps * 1e6, tot / steps * 1e6))
    n = n0
    for label, fn in (("h2d", lambda: stage[:n].copy_(host[0][:n], non_blocking=True)), ("d2h", lambda: host[1][:n].copy_(out[:n], non_blocking=True)),
                      ("h2d quarter", lambda: stage[:n // 4].copy_(host[0][:n // 4], non_blocking=True))):
        torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): fn()
        b.record(); torch.cuda.synchronize()
        print(label, "%.1f us for %d bytes" % (a.elapsed_time(b) / 20 * 1e3, n * 8))
    # resident single step, host waits
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps):
        eng.step(1); torch.cuda.synchronize()
    print("resident step + sync %.1f us per call" % ((time.perf_counter() - t0) / steps * 1e6))
