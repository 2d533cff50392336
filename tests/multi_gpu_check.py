"""Multi-GPU check, run under torchrun on a GPU box (tests/test_gpu_multi.py launches it from `pytest -m gpu` when the
box has at least 2 GPUs):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Chains sharded over ranks + NCCL all-gather swap rounds must give bit-identical chains to the same job on
one GPU (Philox is keyed by the global chain id; every rank replays the same swap round)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bayesbay_b200 as bb  # noqa: E402
from bayesbay_b200 import _dist, synthetic  # noqa: E402
from bayesbay_b200._engine import Engine, pt_swap  # noqa: E402
from bayesbay_b200.forward import RayTomographyForward  # noqa: E402
from tests import problems  # noqa: E402


def run(spec, C, id0, n_total, temps, gather, seed=99, rounds=4, steps=25):
    eng = Engine(spec, C, device=torch.device("cuda", torch.cuda.current_device()), seed=seed, chain_id0=id0)
    eng.init_states()
    eng.temperature.copy_(torch.as_tensor(temps[id0: id0 + C]))
    for r in range(rounds):
        eng.step(steps)
        g = gather(eng.misfit_logdet_T())
        pt_swap(g, seed, r, id0, C, out=eng.temperature)
    k, sites, vals = eng.gather_space(0)
    return k.cpu(), sites.cpu(), vals.cpu(), eng.sigma[0][0].cpu(), eng.temperature.cpu(), eng.n_accepted.cpu()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    spec = problems.spec_from_ingredients(problems.tomo_small())
    n_total = 64 * world
    temps = np.repeat(np.geomspace(1.0, 5.0, 8), n_total // 8)
    id0, C = _dist.shard(n_total)
    mine = run(spec, C, id0, n_total, temps, _dist.all_gather)
    ok = True
    if rank == 0:
        full = run(spec, n_total, 0, n_total, temps, lambda t: t.clone())
    # compare every rank's shard with the single-GPU job (broadcast the full result from rank 0)
    for i in range(6):
        if rank == 0:
            ref = full[i].cuda()
            shape = torch.tensor(list(ref.shape) + [0] * (4 - ref.dim()), device="cuda")
        else:
            shape = torch.zeros(4, dtype=torch.int64, device="cuda")
        dist.broadcast(shape, 0)
        dims = [int(x) for x in shape.tolist() if x > 0]
        if rank != 0:
            ref = torch.empty(dims, dtype=mine[i].dtype, device="cuda")
        dist.broadcast(ref, 0)
        sl = ref[..., id0: id0 + C].cpu()
        same = torch.equal(sl, mine[i]) if i != 1 and i != 2 else torch.equal(torch.nan_to_num(sl), torch.nan_to_num(mine[i]))
        ok &= bool(same)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # the drop-in API under torchrun: n_chains is job-wide, each rank owns a shard, PT swaps over NCCL
    ing = problems.tomo_small()
    fwd = RayTomographyForward((ing["indptr"], ing["indices"], ing["data"]), ing["points"], "voronoi", "vel")
    par, ll = synthetic.build_problem(ing, bb, lambda _: fwd)
    inv = bb.BayesianInversion(par, ll, n_chains=16 * world, seed=5)
    inv.run(sampler=bb.samplers.ParallelTempering(temperature_max=4, swap_every=20), n_iterations=100, save_every=10,
            verbose=False)
    t_local = torch.tensor([ch.temperature for ch in inv.chains], dtype=torch.float64, device="cuda")
    t_all = _dist.all_gather(t_local[:, None])[:, 0].cpu().numpy()
    from oracle import chain as OC

    ladder_ok = np.allclose(np.sort(t_all), np.sort(OC.pt_ladder(16 * world, 4, 0.4)))
    ids_ok = [ch.id for ch in inv.chains] == list(range(rank * 16, rank * 16 + 16))
    if rank == 0:
        print(f"MULTI_GPU_CHECK world={world} sharded==single: {bool(flag.item())} ladder_permuted: {ladder_ok} ids: {ids_ok}",
              flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if (flag.item() and ladder_ok and ids_ok) else 1)


if __name__ == "__main__":
    main()
