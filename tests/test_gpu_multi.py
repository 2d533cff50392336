"""Multi-GPU parity, collected by ``pytest -m gpu``: launches tests/multi_gpu_check.py under torchrun on all the GPUs of
the box (2, 4 or 8) -- chains sharded over ranks with NCCL all-gather swap rounds must be bit-identical to the same job
on one GPU, and the drop-in API must shard ``n_chains`` and permute the reference's ladder.  Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_chains_equal_single_gpu_under_torchrun():
    n = torch.cuda.device_count()
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(REPO, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"MULTI_GPU_CHECK world={n} sharded==single: True ladder_permuted: True ids: True" in out.stdout
