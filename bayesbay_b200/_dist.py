"""The few multi-rank helpers of the path (one process per GPU, ``torch.distributed``).

Chains shard across ranks with no data-path collective; the only exchange is the parallel-tempering
swap round, which needs ``(misfit, logdet, T)`` of every chain on every rank (``all_gather``), and the
Philox key, which must be the same on every rank (``broadcast_seed``).  Backend-agnostic: NCCL on the
GPUs, gloo in the CPU tests."""
from __future__ import annotations


def active():
    """``torch.distributed`` module if a multi-rank job is running, else ``None``."""
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None


def world():
    d = active()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard(n_chains: int):
    """(chain_id0, n_local) of this rank's contiguous shard."""
    rank, size = world()
    if n_chains % size:
        raise ValueError(f"n_chains ({n_chains}) must be divisible by the number of ranks ({size})")
    n_local = n_chains // size
    return rank * n_local, n_local


def all_gather(local):
    """[C_local][F] -> [C_total][F] in global chain order (rank-major = chain-id order)."""
    d = active()
    if d is None:
        return local.clone()
    import torch

    out = torch.empty((local.shape[0] * d.get_world_size(),) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    d.all_gather_into_tensor(out, local.contiguous())
    return out


def broadcast_seed(seed: int, device=None) -> int:
    """Rank 0's seed on every rank."""
    d = active()
    if d is None:
        return int(seed)
    import torch

    t = torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)
    d.broadcast(t, src=0)
    return int(t.item())
