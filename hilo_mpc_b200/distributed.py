"""Multi-GPU plumbing: one process per GPU, batch sharding, and the one collective the path needs.

Instances (initial states, filters, estimators) are independent — nothing in the reference couples them
(``mpc.py``, ``mhe.py``, ``kf.py``, ``pf.py`` operate on one instance per object) — so a batch is split into contiguous chunks,
one per rank, and the solve needs no data-path collective.  The only exchange is the all-gather of
Gaussian-process training shards when the training set is distributed over ranks (``n (d+1) 8`` bytes, latency bound),
done once per fit; an optional gather brings the applied inputs / statuses back to rank 0.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def env_ranks() -> Tuple[int, int, int]:
    """``(rank, world_size, local_rank)`` from the torchrun environment (defaults: single process)."""
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init_process_group(backend: Optional[str] = None):
    """Initialise ``torch.distributed`` from the environment; NCCL when CUDA is visible, gloo otherwise."""
    import torch
    import torch.distributed as dist
    rank, world, local_rank = env_ranks()
    if world == 1 or dist.is_initialized():
        return rank, world, local_rank
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29511')
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend, device_id=torch.device('cuda', local_rank))
    else:
        dist.init_process_group(backend)
    return rank, world, local_rank


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous chunk ``[lo, hi)`` of ``total`` instances owned by ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(array, world: Optional[int] = None, rank: Optional[int] = None):
    """This rank's slice of a batch-major array / tensor."""
    r, w, _ = env_ranks()
    world = w if world is None else world
    rank = r if rank is None else rank
    lo, hi = shard_bounds(len(array), world, rank)
    return array[lo:hi]


def all_gather_variable(t):
    """All-gather of per-rank tensors whose leading dimension may differ; returns the concatenation in rank order."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def all_gather_training_shards(X_shard, y_shard):
    """GP training set distributed by observations: ``X_shard`` ``(d, n_r)``, ``y_shard`` ``(1, n_r)`` on every rank ->
    the full ``(d, n)``, ``(1, n)`` on every rank (rank order), after which the fit is replicated."""
    import torch
    Xt = torch.as_tensor(X_shard)
    yt = torch.as_tensor(y_shard).reshape(1, -1)
    packed = torch.cat([Xt, yt], dim=0).T.contiguous()  # observations on the leading axis
    full = all_gather_variable(packed)
    return full[:, :-1].T.contiguous(), full[:, -1:].T.contiguous()


def gather_to_rank0(t):
    """Concatenate per-rank results on rank 0 (``None`` elsewhere)."""
    import torch.distributed as dist
    full = all_gather_variable(t)
    if not dist.is_initialized() or dist.get_rank() == 0:
        return full
    return None


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction used by the benchmark: maximum over ranks."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
