"""Multi-rank check of the fused detection gather (run under torchrun, or with RANK/WORLD_SIZE=0/1 on one GPU):
every rank decodes ITS sweeps with gather=(PeerGather, slot, sweep_offset); after the device-side barrier each rank
must hold, for every rank, exactly the rows pack_rows() would have produced there."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "range-view-3d-detection_b200")]
from rv3d.distributed import PeerGather, gather_detections, pack_rows  # noqa: E402
from rv3d.nn.decoders.range_decoder import RangeDecoder  # noqa: E402
from tests import synth  # noqa: E402
from tests.util import PP, SBR, ms_outputs, to_dev  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B, C = 3, 3
    dec = RangeDecoder(True, True, *SBR)
    tasks = {0: ["a", "b", "c"]}
    pp = dict(PP, nms_mode="HARD")
    peer = PeerGather(B * C * pp["num_post_nms"], dev)
    for step in range(5):                                  # alternating slots, different data every step
        head = synth.make_head_outputs(B, C, 16, 256, seed=100 * step + rank, n_objects=8)
        if step == 3 and rank == world - 1:
            head["logits"].fill_(-20.0)                    # one rank without any detection
        out = dec.decode(ms_outputs(to_dev(head, dev)), pp, tasks, gather=(peer, step, rank * B))
        peer.arrive_and_wait()
        got = peer.unpack(step)
        ref = gather_detections(pack_rows(*out, batch_offset=rank * B))   # the NCCL form of the same gather
        assert got.shape == ref.shape, (got.shape, ref.shape)
        assert torch.equal(got, ref), f"step {step}: rows differ"
        hdr = peer.rows(step)[:, 0, :2].cpu()
        assert torch.equal(hdr[:, 0], hdr[:, 1])           # nothing dropped
        if step == 3:
            assert int(hdr[world - 1, 0]) == 0
    # pipelined form: the barrier runs on a side stream, a step is consumed after the NEXT one has been launched
    torch.cuda.synchronize()
    pending = None
    for step in range(5, 12):
        head = synth.make_head_outputs(B, C, 16, 256, seed=100 * step + rank, n_objects=8)
        peer.begin(step)
        out = dec.decode(ms_outputs(to_dev(head, dev)), pp, tasks, gather=(peer, step, rank * B))
        peer.publish(step)
        ref = gather_detections(pack_rows(*out, batch_offset=rank * B))
        if pending is not None:
            pstep, pref = pending
            peer.wait(pstep)
            assert torch.equal(peer.unpack(pstep), pref), f"pipelined step {pstep}: rows differ"
        pending = (step, ref)
    peer.wait(pending[0])
    assert torch.equal(peer.unpack(pending[0]), pending[1])
    dist.barrier()
    # sequence-flag form: no barrier kernel; slot and step number live in device memory, so the step also replays as a
    # CUDA graph.  Different data every step; one rank goes empty once; consumers wait with wait_published().
    torch.cuda.synchronize()
    peer2 = PeerGather(B * C * pp["num_post_nms"], dev)
    for step in range(6):
        head = synth.make_head_outputs(B, C, 16, 256, seed=1000 + 100 * step + rank, n_objects=8)
        if step == 2 and rank == 0:
            head["logits"].fill_(-20.0)
        det = dec.decode_async(ms_outputs(to_dev(head, dev)), pp, tasks, gather=peer2.flagged(rank * B))
        peer2.wait_published()
        out = det.result()
        torch.cuda.synchronize()
        got = PeerGather.unpack_rows(peer2.rows_published())
        ref = gather_detections(pack_rows(*out, batch_offset=rank * B))
        assert got.shape == ref.shape and torch.equal(got, ref), f"flagged step {step}: rows differ"
    # the same step captured once and replayed on fresh data
    hd = to_dev(synth.make_head_outputs(B, C, 16, 256, seed=5000 + rank, n_objects=8), dev)
    dec.decode_async(ms_outputs(hd), pp, tasks, gather=peer2.flagged(rank * B))        # warm the workspaces
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=side):
        det = dec.decode_async(ms_outputs(hd), pp, tasks, gather=peer2.flagged(rank * B))
    torch.cuda.current_stream().wait_stream(side)
    for step in range(4):
        new = to_dev(synth.make_head_outputs(B, C, 16, 256, seed=6000 + 100 * step + rank, n_objects=8), dev)
        for k in hd:
            hd[k].copy_(new[k])
        g.replay()
        peer2.wait_published()
        torch.cuda.synchronize()
        peer2.sync_steps()
        m = det.wait()
        mine = pack_rows(det.params[:m], det.scores[:m], det.categories[:m], det.batch_index[:m], batch_offset=rank * B)
        got = PeerGather.unpack_rows(peer2.rows_published())
        ref = gather_detections(mine)
        assert got.shape == ref.shape and torch.equal(got, ref), f"graph replay {step}: rows differ"
    dist.barrier()
    if rank == 0:
        print(f"peer gather ok: world {world}, {got.shape[0]} rows in the last step (barrier, pipelined, sequence-flag and "
              f"graph-replay forms)", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
