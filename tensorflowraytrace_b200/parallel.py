"""
Multi-GPU plumbing: one process per GPU, rays sharded, mesh replicated.

The trace hot path shards over rays (independent units): rank r of W traces the contiguous
block [r*N/W, (r+1)*N/W) of the amalgamated source rays against the full (replicated) boundary
set, so the concatenation over ranks of every per-iteration history equals the single-GPU
history.  The forward pass needs no communication.  The backward pass produces a partial
d(loss)/d(parameters) per rank; because the reference's loss is a plain sum over rays
(tfrt/optimizer.py:219-220: `tape.gradient` of a vector = gradient of its sum), one
sum-all-reduce per optimisation step over NCCL (NVLink 5 / NVSwitch) reproduces the single-GPU
gradient up to fp32 summation order.  All parameter gradients of a step travel in ONE flat
bucket (a few hundred KB: latency-bound, so one collective instead of one per surface), and
every rank then applies the identical update, which keeps the replicas bit-identical without
a broadcast.
"""

import os

import torch
import torch.distributed as dist


def shard_range(n, rank, world_size):
    """Contiguous block of `n` units owned by `rank`: sizes differ by at most one, earlier
    ranks take the remainder."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"shard_range: bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE /
    MASTER_ADDR / MASTER_PORT) and select this rank's GPU.  Returns (rank, world_size,
    local_rank); a no-op single-process answer when WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local % max(torch.cuda.device_count(), 1))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def allreduce_sum_(tensors, group=None):
    """In-place sum-all-reduce of a list of tensors as ONE flat bucket.  `None` entries are
    skipped.  Returns the list.  No-op when not distributed."""
    if not is_distributed():
        return tensors
    live = [t for t in tensors if t is not None]
    if not live:
        return tensors
    if len(live) == 1 and live[0].is_contiguous():
        dist.all_reduce(live[0], op=dist.ReduceOp.SUM, group=group)
        return tensors
    flat = torch.cat([t.reshape(-1).to(live[0].dtype) for t in live])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in live:
        n = t.numel()
        t.copy_(flat[off:off + n].reshape(t.shape))
        off += n
    return tensors


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (timing: the step time is the slowest rank's)."""
    if not is_distributed():
        return float(value)
    device = device or (torch.device("cuda", torch.cuda.current_device())
                        if torch.cuda.is_available() else torch.device("cpu"))
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not is_distributed():
        return float(value)
    device = device or (torch.device("cuda", torch.cuda.current_device())
                        if torch.cuda.is_available() else torch.device("cpu"))
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
