"""Multi-GPU sharding of the measurement path (one process per GPU, torch.distributed for the plumbing).

The path shards along two independent axes and has exactly one exchange step:

* by q-vector      -- every rank sees the same beads and owns a contiguous range of wave-vectors (output columns);
                      results are concatenated with one all-gather per bin.
* by configuration -- every rank accumulates its own walker configurations into its device-resident bin;
                      one reduce (sum) of the bin onto rank 0 per bin.

There is no data-path collective inside a measurement.  The functions below are backend-agnostic (NCCL on GPUs,
gloo in the CPU tests): they move torch tensors, the evaluation itself is whatever the caller runs on its shard.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) split of n items; the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world: int) -> list[int]:
    return [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]


def gather_q_shards(local: torch.Tensor, nq: int, group=None) -> torch.Tensor:
    """local: [nq_local, ...] results of this rank's wave-vectors -> [nq, ...] on every rank (rank order = q order).
    Uneven shards are padded to the largest one for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(nq, world)
    width = max(sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def reduce_bins(local_bin: torch.Tensor, count: int, dst: int = 0, group=None, deterministic: bool = False):
    """Sum the per-rank bins (and the number of configurations in them) onto `dst`.

    deterministic=False: one dist.reduce (NCCL over NVLink; summation order chosen by the library, differences are
    ~1e-16 relative).  deterministic=True: gather to dst and add in rank order, bit-reproducible for a fixed world.
    Returns (bin, count) on dst, (None, None) elsewhere."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    n = torch.tensor([count], dtype=torch.int64, device=local_bin.device)
    dist.reduce(n, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if deterministic:
        parts = [torch.empty_like(local_bin) for _ in range(world)] if rank == dst else None
        dist.gather(local_bin, parts, dst=dst, group=group)
        if rank != dst:
            return None, None
        total = parts[0].clone()
        for p in parts[1:]:
            total += p
        return total, int(n.item())
    total = local_bin.clone()
    dist.reduce(total, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank != dst:
        return None, None
    return total, int(n.item())


class DevPtr:
    """CUDA-array-interface view of a device buffer owned by libpimc_b200 (the bin), so that torch / NCCL can
    operate on it in place."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def bins_tensor(ctx, device: torch.device) -> torch.Tensor:
    ptr, count = ctx.bins_device_ptr()
    return torch.as_tensor(DevPtr(ptr, count), device=device)


def finalize_bin(total: np.ndarray, nq: int, M: int, count: int):
    """EstimatorBase::output normalisation (src/estimator.cpp:351): value * norm / numAccumulated, norm = 1/M."""
    ssf = total[:nq] / (M * count)
    isf = total[nq:].reshape(nq, M) / (M * count)
    return ssf, isf
