"""Sweep sharding and the one collective of the path: a variable-length gather of detections.

Sweeps are independent ("Iterate over all batches since they're mutually exclusive",
math/ops/nms.py:209-210), so a batch shards across ranks with no data-path collective.  The reference
"gathers" detections by writing per-sweep feather files from every rank under a FileLock and a
``dist.barrier()`` (nn/arch/detector.py:366-380, 415-421); here it is one all_gather of counts plus one
padded all_gather of packed rows over NCCL (gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

ROW = 13  # [batch_index, category, score, x, y, z, l, w, h, qw, qx, qy, qz]


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of sweeps owned by ``rank`` (blocks differ by at most one sweep)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_rows(params: Tensor, scores: Tensor, categories: Tensor, batch_index: Tensor, batch_offset: int = 0) -> Tensor:
    """(M,10),(M,),(M,),(M,) -> (M,13) float32 rows; ``batch_offset`` turns the rank-local sweep index
    into the global one."""
    m = params.shape[0]
    rows = torch.empty((m, ROW), dtype=torch.float32, device=params.device)
    rows[:, 0] = batch_index.reshape(m).float() + float(batch_offset)
    rows[:, 1] = categories.reshape(m).float()
    rows[:, 2] = scores.reshape(m).float()
    rows[:, 3:] = params.float()
    return rows


def gather_detections(rows: Tensor, group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """All ranks receive every rank's rows, concatenated in rank order (= global sweep order when the
    shards are contiguous blocks).  Payload is tiny (52 B per detection): latency bound."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    count = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap, ROW), dtype=torch.float32, device=rows.device)
    padded[: rows.shape[0]] = rows
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def gather_detections_fixed(rows: Tensor, capacity: int, group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """Latency-optimised form for a steady-state loop: ONE collective, fixed shape, no host read.

    Every rank contributes a ``(capacity + 1, 13)`` block whose row 0 holds its row count; the result is
    the ``(world, capacity + 1, 13)`` stack, still on the device.  ``unpack_fixed`` turns it into the same
    concatenated rows ``gather_detections`` returns (that is where the host learns the counts)."""
    m = rows.shape[0]
    if m > capacity:
        raise ValueError(f"{m} detections exceed the gather capacity {capacity}")
    block = torch.zeros((capacity + 1, ROW), dtype=torch.float32, device=rows.device)
    block[0, 0] = float(m)
    block[1 : m + 1] = rows
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return block[None]
    world = dist.get_world_size(group)
    out = torch.empty((world, capacity + 1, ROW), dtype=torch.float32, device=rows.device)
    if rows.is_cuda:
        dist.all_gather_into_tensor(out, block, group=group)        # NCCL: one contiguous collective
    else:                                                           # gloo (CPU tests) has no _allgather_base
        dist.all_gather(list(out.unbind(0)), block, group=group)
    return out


def unpack_fixed(stacked: Tensor) -> Tensor:
    counts = stacked[:, 0, 0].to(torch.int64).tolist()
    return torch.cat([stacked[r, 1 : c + 1] for r, c in enumerate(counts)], dim=0)
