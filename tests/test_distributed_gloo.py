"""N>1 host logic on CPU: world_size-2 gloo run of the sweep sharding + detection gather."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_sweeps, out):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "range-view-3d-detection_b200"))
    from rv3d.distributed import gather_detections, gather_detections_fixed, pack_rows, shard_bounds, unpack_fixed
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_sweeps, rank, world)
    g = torch.Generator().manual_seed(1234)
    # every rank builds the same global table and keeps only its shard's detections
    per_sweep = torch.randint(0, 6, (n_sweeps,), generator=g)
    bidx = torch.repeat_interleave(torch.arange(n_sweeps), per_sweep)
    table = torch.rand((int(per_sweep.sum()), 12), generator=g)
    mine = (bidx >= lo) & (bidx < hi)
    rows = pack_rows(table[mine][:, 2:], table[mine][:, 1], table[mine][:, 0], bidx[mine] - lo, batch_offset=lo)
    allrows = gather_detections(rows)
    expect = torch.cat([bidx[:, None].float(), table[:, 0:1], table[:, 1:2], table[:, 2:]], 1)
    ok = torch.equal(allrows, expect)
    ok = ok and torch.equal(unpack_fixed(gather_detections_fixed(rows, 64)), expect)
    dist.barrier()
    dist.destroy_process_group()
    out[rank] = bool(ok)


@pytest.mark.parametrize("n_sweeps", [7, 2, 1])
def test_shard_and_gather_world2(n_sweeps):
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_sweeps, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_bounds_cover():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "range-view-3d-detection_b200"))
    from rv3d.distributed import shard_bounds
    for n in (0, 1, 5, 16, 512, 513):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
