"""World-size-2 checks of the multi-rank logic on CPU (gloo): chain sharding, all-gather order, seed
broadcast, and that every rank replaying the swap round from the shared key ends with the same ladder --
so each rank can keep just its own slice (what the NCCL path does on the GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bayesbay_b200 import _dist
    from oracle import chain as OC
    from oracle import rng as OR

    n_total = 12
    id0, n_local = _dist.shard(n_total)
    assert (id0, n_local) == (rank * 6, 6)
    with pytest.raises(ValueError):
        _dist.shard(13)
    # every rank owns the scalars of its own chains only
    rng = np.random.default_rng(5)
    misfit, logdet = rng.uniform(50, 150, n_total), rng.uniform(-20, 20, n_total)
    ladder = OC.pt_ladder(n_total, 5.0, 0.4)
    local = torch.tensor(np.stack((misfit, logdet, ladder), axis=1)[id0: id0 + n_local])
    gathered = _dist.all_gather(local)
    assert gathered.shape == (n_total, 3)
    np.testing.assert_array_equal(gathered.numpy(), np.stack((misfit, logdet, ladder), axis=1))  # chain-id order
    seed = _dist.broadcast_seed(1000 + rank)
    assert seed == 1000  # rank 0's
    # identical sequential replay on every rank from the shared key; keep the local slice
    g = gathered.numpy()
    temps, n_acc = OC.pt_swap_round(g[:, 0], g[:, 1], list(g[:, 2]), OR.PhiloxStream(seed, 0, 7, OR.STREAM_SWAP))
    mine = torch.tensor(temps[id0: id0 + n_local])
    everyone = _dist.all_gather(mine[:, None])[:, 0].numpy()
    np.testing.assert_array_equal(everyone, np.array(temps))           # slices are consistent across ranks
    np.testing.assert_allclose(np.sort(everyone), np.sort(ladder))     # a permutation of the ladder
    np.save(os.path.join(out_dir, f"temps_{rank}.npy"), np.array(temps))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "temps_0.npy"), np.load(tmp_path / "temps_1.npy")
    np.testing.assert_array_equal(a, b)
    # and equal to the single-process result
    from oracle import chain as OC
    from oracle import rng as OR

    rng = np.random.default_rng(5)
    misfit, logdet = rng.uniform(50, 150, 12), rng.uniform(-20, 20, 12)
    ref, _ = OC.pt_swap_round(misfit, logdet, list(OC.pt_ladder(12, 5.0, 0.4)), OR.PhiloxStream(1000, 0, 7, OR.STREAM_SWAP))
    np.testing.assert_array_equal(a, np.array(ref))


def test_single_process_helpers_are_identity():
    from bayesbay_b200 import _dist

    assert _dist.world() == (0, 1) and _dist.shard(10) == (0, 10)
    t = torch.arange(6.0).reshape(3, 2)
    out = _dist.all_gather(t)
    assert torch.equal(out, t) and out.data_ptr() != t.data_ptr()
    assert _dist.broadcast_seed(42) == 42


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference: under torchrun every rank but 0 exits 0 without work or output."""
    import subprocess

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
