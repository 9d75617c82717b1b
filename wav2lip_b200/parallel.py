"""Multi-GPU layout of the hot path: inference batches shard embarrassingly (eval-mode forward has
no cross-sample op anywhere in wav2lip.py:87-125), so each rank owns a contiguous slice of the batch
dimension B — whole T-windows, so the t-major flatten stays local — and there is NO data-path
collective.  torch.distributed is used only for the rendezvous, barriers and the max-over-ranks
timing reduction; it works with NCCL on GPUs and with gloo on CPU (tests/test_dist_cpu.py)."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of `n` items for `rank` (first n % world ranks get +1)."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} rank={rank} world={world}")
    q, r = divmod(n, world)
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def max_over_ranks(value: float, device=None) -> float:
    """All-reduce(MAX) of a python float; a no-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def sharded_forward(forward_fn, inputs, world_gather: bool = True):
    """Run `forward_fn(*local_inputs)` on this rank's slice of dim 0 of every input and, if
    `world_gather`, all-gather the slices back in rank order (convenience for evaluation scripts; the
    throughput path never gathers).  Works for world size 1 without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return forward_fn(*inputs)
    rank, world = dist.get_rank(), dist.get_world_size()
    n = inputs[0].shape[0]
    b, e = shard_range(n, rank, world)
    local = forward_fn(*[x[b:e] for x in inputs])
    if not world_gather or world == 1:
        return local
    sizes = shard_sizes(n, world)
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)
