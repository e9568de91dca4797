"""One-node multi-GPU partitioning of the path (one process per GPU, torch.distributed over NCCL/NVLink):

  * frames are independent through the encoder  -> contiguous frame ranges per rank (`frame_range`),
  * videos are independent through the decoder   -> round-robin videos per rank (`videos_for_rank`),
  * a single video's decoder stream is sequential -> it lives on ONE owner rank; the only data-path exchange is the
    gather of encoded frame tokens ([n, 49, 3584] bf16, 351,232 B per frame) to that rank (`FrameParallelEncoder`),
    sent batch by batch so the owner can start decoding the first frames while later ones are still being encoded
    (`FrameParallelEncoder`: NCCL send/recv), or stored directly into the owner's memory by the producing kernel
    (`PeerStoreEncoder`: symmetric memory over NVLink, no collective).

The reference has no multi-GPU inference mode besides accelerate's layer-wise `device_map='auto'`
(models/modeling_live.py:99), which gives no speed-up; this module is what replaces it.  Backend-agnostic on purpose
(NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def frame_range(n_frames, world, rank):
    """Contiguous split; the first n % world ranks take one extra frame."""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def encoder_frame_range(n_frames, encoders, rank):
    """Frame range of `rank` when only the ranks in `encoders` (ascending) encode; (0, 0) for a rank that does not.  Leaving the
    decoder-owning rank out lets it start decoding the first batch while the others are still encoding: the owner's decoder
    stream is the serial term of the pipeline, so nothing else should sit on its GPU."""
    if rank not in encoders:
        return 0, 0
    return frame_range(n_frames, len(encoders), encoders.index(rank))


def encoder_batches(n_frames, encoders, rank, batch, assignment="contiguous"):
    """[(b0, b1)] frame batches `rank` encodes, in its encoding order.  "contiguous": one frame range per encoder;
    "round_robin": batch b of the video goes to encoder b mod E, so the frames become available IN ORDER at the encoders'
    aggregate rate — what a decoder that consumes the video front to back needs (with contiguous ranges it is fed by the first
    encoder alone until that one's range is finished)."""
    if rank not in encoders:
        return []
    if assignment == "round_robin":
        # `batch` may be a list of batch sizes (a schedule, e.g. small batches first so that the first frames land early);
        # the last size repeats until the video is covered
        sizes = list(batch) if isinstance(batch, (list, tuple)) else [batch]
        every, b0, i = [], 0, 0
        while b0 < n_frames:
            b1 = min(b0 + sizes[min(i, len(sizes) - 1)], n_frames)
            every.append((b0, b1))
            b0, i = b1, i + 1
        return every[encoders.index(rank)::len(encoders)]
    if isinstance(batch, (list, tuple)):
        raise ValueError("a batch-size schedule needs assignment='round_robin'")
    lo, hi = encoder_frame_range(n_frames, encoders, rank)
    return [(b0, min(b0 + batch, hi)) for b0 in range(lo, hi, batch)]


def videos_for_rank(n_videos, world, rank):
    return list(range(rank, n_videos, world))


class FrameParallelEncoder:
    """encode_fn(frames[b0:b1]) -> tokens [(b1-b0)*tokens_per_frame, hidden].  Every rank encodes its frame range in
    batches; non-owner ranks isend each finished batch, the owner irecvs straight into the frame-ordered output."""

    def __init__(self, encode_fn, tokens_per_frame, hidden, dtype=torch.bfloat16, device="cuda", owner=0, batch=32, group=None,
                 encoders=None):
        self.encode_fn, self.tpf, self.hidden, self.dtype, self.device = encode_fn, tokens_per_frame, hidden, dtype, device
        self.owner, self.batch, self.group = owner, batch, group
        self.encoders = sorted(encoders) if encoders is not None else None      # ranks that encode (default: all)

    def encode(self, n_frames, local_frames):
        """local_frames: this rank's slice (frame_range) of the video.  Returns (tokens, ready) on the owner — `ready[i]`
        is a callable that blocks until frame i's tokens have arrived — and (None, None) elsewhere."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        encoders = self.encoders if self.encoders is not None else list(range(world))
        lo, hi = encoder_frame_range(n_frames, encoders, rank)
        assert len(local_frames) == hi - lo, (len(local_frames), lo, hi)
        is_owner = rank == self.owner
        out, recvs = None, {}
        if is_owner:
            out = torch.empty(n_frames * self.tpf, self.hidden, dtype=self.dtype, device=self.device)
            for src in encoders:
                if src == rank:
                    continue
                s_lo, s_hi = encoder_frame_range(n_frames, encoders, src)
                for b0 in range(s_lo, s_hi, self.batch):
                    b1 = min(b0 + self.batch, s_hi)
                    recvs[(b0, b1)] = dist.irecv(out[b0 * self.tpf:b1 * self.tpf], src=src, group=self.group)
        sends = []
        for b0 in range(lo, hi, self.batch):
            b1 = min(b0 + self.batch, hi)
            tok = self.encode_fn(local_frames[b0 - lo:b1 - lo])
            assert tok.shape == ((b1 - b0) * self.tpf, self.hidden), tok.shape
            if is_owner:
                out[b0 * self.tpf:b1 * self.tpf].copy_(tok)
            else:
                sends.append((dist.isend(tok.contiguous(), dst=self.owner, group=self.group), tok))
        for req, _ in sends:
            req.wait()
        if not is_owner:
            return None, None

        done = set()

        def ready_fn(i):
            def wait():
                for (b0, b1), req in recvs.items():
                    if b0 <= i < b1 and (b0, b1) not in done:   # a Work object must be waited on only once
                        req.wait()
                        done.add((b0, b1))
            return wait
        return out, [ready_fn(i) for i in range(n_frames)]

    @staticmethod
    def wait_all(ready):
        for r in ready or []:
            r()


class PeerStoreEncoder:
    """Same contract as FrameParallelEncoder with no collective call on the data path: the frame-token buffer is
    symmetric memory (torch.distributed._symmetric_memory: one allocation per rank, every rank maps every peer's), and
    each rank's projector + pooling kernel writes its frames STRAIGHT INTO THE OWNER'S HBM through that mapping — the
    kernel's epilogue is the NVLink transfer, there is no staging copy and no send/recv.  A per-batch signal (symmetric
    signal pad, release/acquire at system scope, enqueued on the producing stream) tells the owner which frames landed.

    embed_into(frames[b0:b1], dst[(b1-b0)*tokens_per_frame, hidden]) must write the tokens of those frames into dst."""

    def __init__(self, embed_into, tokens_per_frame, hidden, max_frames, device, owner=0, batch=32, group=None, symm=None,
                 dtype=torch.bfloat16, encoders=None, assignment="contiguous"):
        if symm is None:                   # injectable: the CPU tests drive the same control flow through a gloo-backed stand-in
            import torch.distributed._symmetric_memory as symm
        self.embed_into, self.tpf, self.hidden, self.device = embed_into, tokens_per_frame, hidden, device
        self.owner, self.batch, self.max_frames = owner, batch, max_frames
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.encoders = sorted(encoders) if encoders is not None else list(range(self.world))   # ranks that encode
        self.assignment = assignment       # see encoder_batches
        self.buf = symm.empty(max_frames * tokens_per_frame, hidden, dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        # the owner's buffer as seen from this rank (a peer mapping unless this rank is the owner)
        self.dst = self.buf if self.rank == owner else self.hdl.get_buffer(owner, tuple(self.buf.shape), dtype)
        self._pending = {}                 # owner: signals of the previous encode() that nobody has waited on yet

    N_CHANNELS = 15          # signal channels 1..15 (0 is the barrier's); the signal pad holds world x channels words

    def _channel(self, n):
        if n >= self.N_CHANNELS:
            raise ValueError(f"PeerStoreEncoder: more than {self.N_CHANNELS} batches per rank; raise `batch`")
        return 1 + n

    def encode(self, n_frames, local_frames):
        """Returns (tokens, ready) on the owner (tokens is a view of the symmetric buffer, valid until the next encode;
        ready[i]() makes the current stream wait until frame i has landed) and (None, None) elsewhere."""
        assert n_frames <= self.max_frames
        mine = encoder_batches(n_frames, self.encoders, self.rank, self.batch, self.assignment)
        assert len(local_frames) == sum(b1 - b0 for b0, b1 in mine), (len(local_frames), mine)   # this rank's batches, concatenated
        # Signals of the previous video that the owner never consumed (it stopped early, or never called its ready[i]) would
        # satisfy THIS video's waits before the data has landed, and a producer's put_signal blocks on a channel that is
        # still set: consume them first.  Every producer always posts all of its signals, so these waits terminate.
        for src, batches in self._pending.items():
            for _, _, ch in batches:
                self.hdl.wait_signal(src, channel=ch)
        self._pending = {}
        # nobody may overwrite the owner's buffer while it is still decoding the previous video
        self.hdl.barrier(channel=0)
        off = 0
        own_events = []                     # owner that also encodes: one event per own batch (it may encode on a side stream)
        for n, (b0, b1) in enumerate(mine):
            self.embed_into(local_frames[off:off + b1 - b0], self.dst[b0 * self.tpf:b1 * self.tpf])
            off += b1 - b0
            if self.rank == self.owner and torch.device(self.device).type == "cuda":
                ev = torch.cuda.Event()
                ev.record()
                own_events.append((b0, b1, ev))
            if self.rank != self.owner:
                # stream-ordered after the kernel that stored the batch.  One channel per batch: a signal is a binary
                # semaphore, and reusing one channel would block this rank's stream until the owner had consumed the
                # previous batch (the owner consumes in frame order, so ranks would serialise behind each other).
                self.hdl.put_signal(self.owner, channel=self._channel(n))
        if self.rank != self.owner:
            return None, None
        pending = {}                                           # src rank -> list of its batches, in sending order
        for src in self.encoders:
            if src != self.owner:
                pending[src] = [(b0, b1, self._channel(n))
                                for n, (b0, b1) in enumerate(encoder_batches(n_frames, self.encoders, src, self.batch, self.assignment))]
        self._pending = pending

        def ready_fn(i):
            def wait():
                for b0, b1, ev in own_events:
                    if b0 <= i < b1:
                        torch.cuda.current_stream().wait_event(ev)
                for src, batches in pending.items():
                    if batches and any(b0 <= i < b1 for b0, b1, _ in batches):
                        while batches:                        # consume this source's signals up to the batch holding frame i
                            b0, b1, ch = batches.pop(0)
                            self.hdl.wait_signal(src, channel=ch)
                            if b0 <= i < b1:
                                break
            return wait
        return self.buf[:n_frames * self.tpf], [ready_fn(i) for i in range(n_frames)]

    wait_all = staticmethod(FrameParallelEncoder.wait_all)


def layer_ranges(n_layers, n_stages):
    """Contiguous decoder-layer ranges of a layer pipeline, sizes differing by at most one (earlier stages take the extra)."""
    base, extra = divmod(n_layers, n_stages)
    out, l0 = [], 0
    for s in range(n_stages):
        n = base + (1 if s < extra else 0)
        out.append((l0, l0 + n))
        l0 += n
    return out


class LayerPipeline:
    """One video's decoder split BY LAYERS over `stage_ranks` (ascending pipeline order, one stage per GPU).

    A video's decoder stream is sequential in time — pass p+1 attends over pass p's keys — but only layer by layer: layer l of
    pass p+1 needs layer l's K/V of pass p, not pass p's final output (no token of the video depends on a generated one while
    frames are being scored).  So stage s runs its layers for pass p while stage s-1 already runs pass p+1: the owner's serial
    term is divided by the number of stages, at the price of one point-to-point hand-over of the fp32 residual
    stream [rows, hidden] per pass and stage boundary (28 MB for a 40-frame pass: ~40 us over NVLink).

    run(passes, stage_fn): `passes` is the same list on every stage rank (each entry a dict with 'rows' = tokens of the pass);
    stage_fn(p, desc, resid_in) returns the residual stream to hand on (every stage but the last; resid_in is None on the
    first) or the stage's result (last stage).  Returns the list of results on the last stage, None elsewhere."""

    def __init__(self, stage_ranks, hidden, device, dtype=torch.float32, group=None):
        self.stage_ranks = list(stage_ranks)
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.hidden, self.device, self.dtype = hidden, device, dtype
        self.index = self.stage_ranks.index(self.rank) if self.rank in self.stage_ranks else None

    @property
    def n_stages(self):
        return len(self.stage_ranks)

    @property
    def is_first(self):
        return self.index == 0

    @property
    def is_last(self):
        return self.index == len(self.stage_ranks) - 1

    def run(self, passes, stage_fn):
        if self.index is None:
            return None
        prev = self.stage_ranks[self.index - 1] if not self.is_first else None
        nxt = self.stage_ranks[self.index + 1] if not self.is_last else None
        results, inflight = [], []
        for p, desc in enumerate(passes):
            rin = None
            if prev is not None:
                rin = torch.empty(desc["rows"], self.hidden, dtype=self.dtype, device=self.device)
                dist.irecv(rin, src=prev, group=self.group).wait()      # (NCCL: the current stream waits, not the host)
            out = stage_fn(p, desc, rin)
            if nxt is not None:
                # asynchronous: this stage goes on to pass p + 1 while the hand-over of pass p is in flight; the tensors stay
                # referenced until the send has been waited for
                inflight.append((dist.isend(out, dst=nxt, group=self.group), out))
                if len(inflight) > 2:
                    inflight.pop(0)[0].wait()
            else:
                results.append(out)
        for req, _ in inflight:
            req.wait()
        return results if self.is_last else None


def gather_results(obj, dst=0, group=None):
    """Per-rank Python results (score traces of the rank's videos) -> list on `dst` (a few KB; config 4's only exchange)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(obj, bucket, dst=dst, group=group)
    return bucket
