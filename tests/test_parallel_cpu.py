"""World-size-2 gloo tests of the multi-GPU host logic (no GPU): frame/video partitioning and the chunked gather of
frame tokens to the decoder-owning rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmduet_b200.parallel import encoder_frame_range, FrameParallelEncoder, PeerStoreEncoder, frame_range, gather_results, videos_for_rank


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


# ---------------------------------------------------------------------------------------------------------------------
# PeerStoreEncoder: same control flow as on the GPUs (barrier, one signal channel per batch, in-order consumption), with a
# gloo-backed stand-in for torch's CUDA symmetric memory: "peer stores" land in a local staging tensor and the signal
# carries them (send on put_signal, recv into the owner's buffer on wait_signal).
# ---------------------------------------------------------------------------------------------------------------------
class _GlooSymm:
    class Handle:
        def __init__(self, buf, group):
            self.buf, self.group, self.rank = buf, group, dist.get_rank(group)
            self.staging = torch.full_like(buf, float("nan"))
            self.sent = {}          # channel -> rows already shipped by this rank
            self.log = []

        def get_buffer(self, rank, sizes, dtype):
            assert tuple(sizes) == tuple(self.buf.shape) and dtype == self.buf.dtype
            return self.staging    # what this rank writes "into the owner's memory"

        def barrier(self, channel=0):
            dist.barrier(self.group)
            self.sent = {}                       # a new encode() starts: every channel is free again
            self.staging.fill_(float("nan"))

        def put_signal(self, dst_rank, channel=0):
            assert channel not in self.sent, "a signal channel is a binary semaphore: one batch per channel and encode"
            rows = (~torch.isnan(self.staging[:, 0])).nonzero().flatten()
            new = rows[~torch.isin(rows, torch.tensor(sorted(r for v in self.sent.values() for r in v), dtype=torch.long))]
            self.sent[channel] = new.tolist()
            meta = torch.tensor([channel, int(new[0]), int(new[-1]) + 1])
            dist.send(meta, dst_rank, group=self.group)
            dist.send(self.staging[int(new[0]):int(new[-1]) + 1].contiguous(), dst_rank, group=self.group)

        def wait_signal(self, src_rank, channel=0):
            meta = torch.zeros(3, dtype=torch.long)
            dist.recv(meta, src_rank, group=self.group)
            assert int(meta[0]) == channel, (int(meta[0]), channel)      # signals of one source are consumed in sending order
            dist.recv(self.buf[int(meta[1]):int(meta[2])], src_rank, group=self.group)
            self.log.append((src_rank, channel))

    @staticmethod
    def empty(*size, dtype=None, device=None):
        return torch.full(size, float("nan"), dtype=dtype, device=device)

    @staticmethod
    def rendezvous(tensor, group):
        return _GlooSymm.Handle(tensor, group)


def test_encoder_batches_assignments():
    from mmduet_b200.parallel import encoder_batches
    enc = [1, 2, 3]
    rr = [encoder_batches(10, enc, r, 2, "round_robin") for r in range(4)]
    assert rr[0] == [] and rr[1] == [(0, 2), (6, 8)] and rr[2] == [(2, 4), (8, 10)] and rr[3] == [(4, 6)]
    sched = [encoder_batches(23, [2, 3], r, [2, 2, 4, 8], "round_robin") for r in range(4)]   # sizes 2 2 4 8 then 8s (7 left)
    assert sched[0] == sched[1] == [] and sched[2] == [(0, 2), (4, 8), (16, 23)] and sched[3] == [(2, 4), (8, 16)]
    cont = [encoder_batches(10, enc, r, 2) for r in range(4)]
    assert cont[0] == [] and sorted(b for c in cont for b in c) == sorted(set(b for c in cont for b in c))
    assert sum(b1 - b0 for c in cont for b0, b1 in c) == 10 and sum(b1 - b0 for c in rr for b0, b1 in c) == 10


def _peer_worker(rank, world, port, n_frames, owner, q, owner_encodes=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tpf, hidden = 3, 4
        encoders = list(range(world)) if owner_encodes else [r for r in range(world) if r != owner]
        lo, hi = encoder_frame_range(n_frames, encoders, rank)
        frames = torch.arange(lo, hi, dtype=torch.float32)

        def embed_into(fr, dst):
            dst.copy_((fr[:, None] * 10 + torch.arange(tpf)[None, :]).reshape(-1, 1).expand(-1, hidden))

        enc = PeerStoreEncoder(embed_into, tpf, hidden, max_frames=n_frames, device="cpu", owner=owner, batch=2, symm=_GlooSymm,
                               dtype=torch.float32, encoders=encoders)
        out, ready = enc.encode(n_frames, frames)
        if rank == owner:
            PeerStoreEncoder.wait_all(ready)
            want = (torch.arange(n_frames)[:, None] * 10 + torch.arange(tpf)[None, :]).reshape(-1, 1).expand(-1, hidden).float()
            src = 1 - owner
            n_batches = (encoder_frame_range(n_frames, encoders, src)[1] - encoder_frame_range(n_frames, encoders, src)[0] + 1) // 2
            ok = torch.equal(out, want) and enc.hdl.log == [(src, 1 + b) for b in range(n_batches)]
        else:
            ok = out is None and ready is None
        # second video: the owner walks away without waiting for anything (early stop); third video: the leftover signals
        # must have been consumed before the new ones are trusted (no stale "frame has landed")
        enc.encode(n_frames, frames)
        out, ready = enc.encode(n_frames, frames)
        if rank == owner:
            PeerStoreEncoder.wait_all(ready)
            ok = ok and torch.equal(out, want) and enc.hdl.log == [(src, 1 + b) for b in range(n_batches)] * 3
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,owner,owner_encodes", [(7, 0, True), (12, 1, True), (9, 0, False)])
def test_peer_store_encoder_world2_control_flow(n_frames, owner, owner_encodes):
    """owner_encodes=False: the decoder-owning rank only receives (what bench.py's configs[2] measurement does for N > 1)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, n_frames, owner, q, owner_encodes)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p_ in procs:
        p_.join(timeout=60)
    assert res == {0: True, 1: True}


# ---------------------------------------------------------------------------------------------------------------------
# LayerPipeline: hand-over schedule of a 3-stage decoder pipeline over gloo (the stages' arithmetic is a stand-in)
# ---------------------------------------------------------------------------------------------------------------------
def test_layer_ranges():
    from mmduet_b200.parallel import layer_ranges
    assert layer_ranges(28, 4) == [(0, 7), (7, 14), (14, 21), (21, 28)]
    assert layer_ranges(28, 3) == [(0, 10), (10, 19), (19, 28)]
    assert layer_ranges(28, 1) == [(0, 28)]
    for n in (1, 2, 5, 28):
        r = layer_ranges(28, n)
        assert r[0][0] == 0 and r[-1][1] == 28 and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _pipe_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmduet_b200.parallel import LayerPipeline
        stage_ranks = [0, 2, 1]                                        # pipeline order need not be rank order; rank 3 is no stage
        pipe = LayerPipeline(stage_ranks, hidden=4, device="cpu", dtype=torch.float32)
        passes = [{"rows": 3 + (p % 2), "p": p} for p in range(7)]
        log = []

        def stage_fn(p, desc, rin):
            log.append(p)
            if pipe.is_first:
                x = torch.full((desc["rows"], 4), float(p))
            else:
                assert rin.shape == (desc["rows"], 4)
                x = rin
            y = x * 2 + pipe.index                                     # stage k: y = 2x + k
            return y if not pipe.is_last else y.sum().item()

        res = pipe.run(passes, stage_fn)
        ok = True
        if rank == 3:
            ok = res is None and log == []
        else:
            ok = log == list(range(7))
            if pipe.is_last:
                # ((2p + 0) * 2 + 1) * 2 + 2 = 8p + 4 per element
                want = [(8 * p + 4) * 4 * (3 + p % 2) for p in range(7)]
                ok = ok and res == want
            else:
                ok = ok and res is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_layer_pipeline_hand_over_gloo():
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipe_worker, args=(r, world, port, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p_ in procs:
        p_.join(timeout=60)
    assert res == {0: True, 1: True, 2: True, 3: True}


def _peer_rr_worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmduet_b200.parallel import encoder_batches
        tpf, hidden, owner, encoders = 3, 4, 0, [1, 2]
        mine = encoder_batches(n_frames, encoders, rank, 2, "round_robin")
        frames = torch.tensor([float(f) for b0, b1 in mine for f in range(b0, b1)])

        def embed_into(fr, dst):
            dst.copy_((fr[:, None] * 10 + torch.arange(tpf)[None, :]).reshape(-1, 1).expand(-1, hidden))

        enc = PeerStoreEncoder(embed_into, tpf, hidden, max_frames=n_frames, device="cpu", owner=owner, batch=2, symm=_GlooSymm,
                               dtype=torch.float32, encoders=encoders, assignment="round_robin")
        ok = True
        for _ in range(2):
            out, ready = enc.encode(n_frames, frames)
            if rank == owner:
                want = (torch.arange(n_frames)[:, None] * 10 + torch.arange(tpf)[None, :]).reshape(-1, 1).expand(-1, hidden).float()
                for i in range(n_frames):                              # front to back, as a decoder consumes the video
                    ready[i]()
                    ok = ok and torch.equal(out[i * tpf:(i + 1) * tpf], want[i * tpf:(i + 1) * tpf])
            else:
                ok = ok and out is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_peer_store_round_robin_batches_world3():
    """Owner 0 only receives; encoders 1 and 2 take the 2-frame batches alternately; the owner reads the video front to back."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_rr_worker, args=(r, 3, port, 11, q)) for r in range(3)]
    for p_ in procs:
        p_.start()
    res = dict(q.get(timeout=120) for _ in range(3))
    for p_ in procs:
        p_.join(timeout=60)
    assert res == {0: True, 1: True, 2: True}
