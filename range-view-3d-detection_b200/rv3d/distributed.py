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


class PeerGather:
    """The gather of detections as peer-memory stores fused into the NMS pack kernel (``rv3d_nms`` with ``peer_world``).

    Every rank owns a ``(slots, world, capacity + 1, 16)`` float32 buffer in symmetric memory (torch's symmetric-memory
    allocator maps all ranks' buffers into each process: plumbing only).  Rank r's pack kernel writes its detections
    into slot ``r`` of EVERY rank's buffer through the NVLink-mapped pointers; ``arrive_and_wait`` is a device-side
    barrier over the ranks on the current stream (no host synchronisation), after which ``rows`` / ``unpack`` read the
    complete gather locally.  Two slots alternate between consecutive steps so a step's writers never race the previous
    step's readers.  Raises if symmetric memory is unavailable (callers fall back to ``gather_detections_fixed``)."""

    ROW16 = 16

    def __init__(self, capacity: int, device: torch.device, group: Optional[dist.ProcessGroup] = None, slots: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("PeerGather handles up to 8 ranks (one NVSwitch domain)")
        self.capacity, self.slots = int(capacity), int(slots)
        self.buf = symm_mem.empty((self.slots, self.world, self.capacity + 1, self.ROW16), dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.slot_bytes = self.world * (self.capacity + 1) * self.ROW16 * 4
        # sequence-flag protocol: this rank's count of published steps, advanced by the pack kernel itself
        self.seq = torch.zeros(1, dtype=torch.int32, device=device)
        self.steps = 0
        self.hdl.barrier()

    def slot_ptrs(self, slot: int):
        return [p + (slot % self.slots) * self.slot_bytes for p in self.ptrs]

    def write_empty(self, slot: int) -> None:
        """This rank has no detections this step: publish a zero header to every rank."""
        for q in range(self.world):
            peer = self.hdl.get_buffer(q, (self.slots, self.world, self.capacity + 1, self.ROW16), torch.float32)
            peer[slot % self.slots, self.rank, 0, :4] = 0.0

    def arrive_and_wait(self) -> None:
        """Synchronous form: barrier on the current stream right after the writes."""
        self.hdl.barrier()

    # ---- pipelined form: the barrier runs on a side stream, so a rank may run one step ahead of the slowest one ----
    #   begin(slot)    before the NMS call that writes `slot`: wait until the barrier of the PREVIOUS step is through
    #                  (then every rank has finished reading this slot's previous contents, two steps back)
    #   publish(slot)  after the NMS call: the side stream waits for the writes, runs the barrier, marks the slot gathered
    #   wait(slot)     a consumer's stream waits for the slot to be gathered;  side_stream runs consumers off-path
    def _lazy_async(self):
        if not hasattr(self, "side_stream"):
            self.side_stream = torch.cuda.Stream(self.buf.device)
            self._written = [torch.cuda.Event() for _ in range(self.slots)]
            self._done = [torch.cuda.Event() for _ in range(self.slots)]
            self._published = [False] * self.slots

    def begin(self, slot: int) -> None:
        self._lazy_async()
        prev = (slot - 1) % self.slots
        if self._published[prev]:
            torch.cuda.current_stream(self.buf.device).wait_event(self._done[prev])

    def publish(self, slot: int) -> None:
        self._lazy_async()
        s = slot % self.slots
        self._written[s].record(torch.cuda.current_stream(self.buf.device))
        with torch.cuda.stream(self.side_stream):
            self.side_stream.wait_event(self._written[s])
            self.hdl.barrier()
            self._done[s].record(self.side_stream)
        self._published[s] = True

    def wait(self, slot: int) -> None:
        self._lazy_async()
        if self._published[slot % self.slots]:
            torch.cuda.current_stream(self.buf.device).wait_event(self._done[slot % self.slots])

    # ---- sequence-flag form (slots == 2): no barrier kernel, no host-side slot bookkeeping, graph-replayable ----
    #   decode_async(..., gather=self.flagged(sweep_offset))   the pack kernel writes step k = seq + 1 into slot k & 1 of
    #                                                          every rank and publishes k in the slot headers
    #   wait_published()   enqueue a wait until every rank's step `seq` has arrived in THIS rank's buffer.  Call it before
    #                      the next step's NMS (it then overlaps the next rasterize + decode) and before reading rows.
    def flagged(self, sweep_offset: int = 0):
        if self.slots != 2:
            raise ValueError("the sequence-flag protocol alternates between exactly two slots")
        self.steps += 1
        return (self, 0, int(sweep_offset), True)

    def wait_published(self) -> None:
        from . import _native as N
        from ._util import stream_ptr
        import ctypes as C
        N.check(N.lib().rv3d_peer_wait(C.c_void_p(self.buf.data_ptr()), self.slot_bytes // 4, self.world, self.capacity,
                                       C.c_void_p(self.seq.data_ptr()), stream_ptr(self.buf.device)), "rv3d_peer_wait")

    def sync_steps(self) -> int:
        """Re-read the device-side step counter (after CUDA-graph replays, which advance it without the host noticing)."""
        self.steps = int(self.seq.item())
        return self.steps

    def rows_published(self) -> Tensor:
        """Rows of the last published step (host-side step count; valid after ``wait_published``)."""
        return self.buf[self.steps & 1]

    def rows(self, slot: int) -> Tensor:
        """(world, capacity + 1, 16): row 0 of each rank's block is [rows written, rows kept, 0, 0]."""
        return self.buf[slot % self.slots]

    def unpack(self, slot: int) -> Tensor:
        """-> (M, 13) rows [sweep, class, score, params(10)] of all ranks in rank order (host reads the counts)."""
        return self.unpack_rows(self.rows(slot))

    @staticmethod
    def unpack_rows(blk: Tensor) -> Tensor:
        head = blk[:, 0, :2].cpu()
        counts, kept = head[:, 0].to(torch.int64).tolist(), head[:, 1].to(torch.int64).tolist()
        if any(k > c for c, k in zip(counts, kept)):
            raise RuntimeError(f"PeerGather: a rank kept {max(kept)} detections but the gather buffer holds "
                               f"{blk.shape[1] - 1} rows per rank; raise `capacity`")
        cols = [0, 1, 2] + list(range(4, 14))
        return torch.cat([blk[r, 1 : c + 1][:, cols] for r, c in enumerate(counts)], dim=0)
