"""World-size-2 gloo tests of the multi-GPU host logic (no GPU): frame/video partitioning and the chunked gather of
frame tokens to the decoder-owning rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmduet_b200.parallel import FrameParallelEncoder, frame_range, gather_results, videos_for_rank


def test_partitions_cover_everything_once():
    for n in (0, 1, 7, 120, 600, 601):
        for world in (1, 2, 4, 8):
            spans = [frame_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            vids = sorted(v for r in range(world) for v in videos_for_rank(n, world, r))
            assert vids == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, owner, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tpf, hidden = 3, 4
        lo, hi = frame_range(n_frames, world, rank)
        frames = torch.arange(lo, hi, dtype=torch.float32)          # a "frame" is just its global index

        def encode(fr):                                                # token j of frame i = i*10 + j in every channel
            return (fr[:, None] * 10 + torch.arange(tpf)[None, :]).reshape(-1, 1).expand(-1, hidden).contiguous()

        enc = FrameParallelEncoder(encode, tpf, hidden, dtype=torch.float32, device="cpu", owner=owner, batch=2)
        out, ready = enc.encode(n_frames, frames)
        if rank == owner:
            ready[n_frames - 1]()                                      # waiting on one frame only needs its batch
            FrameParallelEncoder.wait_all(ready)
            want = (torch.arange(n_frames)[:, None] * 10 + torch.arange(tpf)[None, :]).reshape(-1, 1).expand(-1, hidden).float()
            ok = torch.equal(out, want)
        else:
            ok = out is None and ready is None
        res = gather_results({"rank": rank, "videos": videos_for_rank(5, world, rank)}, dst=0)
        if rank == 0:
            ok = ok and [r["rank"] for r in res] == list(range(world)) and sorted(v for r in res for v in r["videos"]) == list(range(5))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,owner", [(7, 0), (9, 1), (2, 0), (1, 1)])
def test_frame_parallel_gather_gloo(n_frames, owner):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, owner, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}
