"""Developer probe (torchrun, N >= 1): device time of the pieces of a tempered window -- 20 iterations, the swap rows
kernel, the NCCL all-gather, the swap kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bayesbay_b200 import synthetic
from bayesbay_b200._engine import Engine, pt_swap
from tests import problems

world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
C = 1024
spec = problems.spec_from_ingredients(synthetic.tomography_2d())
eng = Engine(spec, C, device=dev, seed=12345, chain_id0=rank * C)
eng.init_states()
ladder = np.tile(np.geomspace(1.0, 5.0, 16), world * C // 16)
eng.temperature.copy_(torch.as_tensor(ladder[rank * C:(rank + 1) * C]))
eng.step(300)
local = torch.empty((C, 3), dtype=torch.float64, device=dev)
gathered = torch.empty((world * C, 3), dtype=torch.float64, device=dev) if world > 1 else local
def ev(): return torch.cuda.Event(enable_timing=True)
names = ["step(20)", "rows", "all_gather", "pt_swap"]
acc = np.zeros(4); n = 0
for r in range(30):
    e = [ev() for _ in range(5)]
    e[0].record(); eng.step(20)
    e[1].record(); eng.misfit_logdet_T(local)
    e[2].record()
    if world > 1: dist.all_gather_into_tensor(gathered, local)
    e[3].record(); pt_swap(gathered, 777, r, rank * C, C, out=eng.temperature)
    e[4].record(); torch.cuda.synchronize()
    if r >= 5:
        acc += [e[i].elapsed_time(e[i + 1]) for i in range(4)]; n += 1
if rank == 0:
    print("world", world, {k: round(v / n, 4) for k, v in zip(names, acc)}, "ms", flush=True)
e0, e1 = ev(), ev()
torch.cuda.synchronize(); 
if world > 1: dist.barrier()
e0.record()
for r in range(30):
    eng.step(20); eng.misfit_logdet_T(local)
    if world > 1: dist.all_gather_into_tensor(gathered, local)
    pt_swap(gathered, 777, 100 + r, rank * C, C, out=eng.temperature)
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print("30 windows back to back: %.4f ms per window (%.4f ms/step)" % (e0.elapsed_time(e1) / 30, e0.elapsed_time(e1) / 600), flush=True)
if world > 1: dist.destroy_process_group()
# host-side time of each call in the back-to-back loop
import time
if world == 1:
    torch.cuda.synchronize()
    hs = np.zeros(3)
    t_all0 = time.perf_counter()
    for r in range(30):
        t0 = time.perf_counter(); eng.step(20)
        t1 = time.perf_counter(); eng.misfit_logdet_T(local)
        t2 = time.perf_counter(); pt_swap(gathered, 777, 200 + r, rank * C, C, out=eng.temperature)
        t3 = time.perf_counter()
        hs += [t1 - t0, t2 - t1, t3 - t2]
    t_host = time.perf_counter() - t_all0
    torch.cuda.synchronize()
    t_tot = time.perf_counter() - t_all0
    print("host ms per window: step(20) %.3f rows %.3f pt_swap %.3f | enqueue total %.3f, with final sync %.3f" % (
        *(hs / 30 * 1e3), t_host / 30 * 1e3, t_tot / 30 * 1e3), flush=True)
    # the same loop with step(20) only
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for r in range(30): eng.step(20)
    torch.cuda.synchronize(); print("step(20) x 30 only: %.3f ms per window" % ((time.perf_counter() - t0) / 30 * 1e3), flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for r in range(30): eng.step(20); eng.misfit_logdet_T(local)
    torch.cuda.synchronize(); print("step(20) + rows x 30: %.3f ms per window" % ((time.perf_counter() - t0) / 30 * 1e3), flush=True)
