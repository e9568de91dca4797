"""One-node multi-GPU partitioning of the path (one process per GPU, torch.distributed over NCCL/NVLink):

  * frames are independent through the encoder  -> contiguous frame ranges per rank (`frame_range`),
  * videos are independent through the decoder   -> round-robin videos per rank (`videos_for_rank`),
  * a single video's decoder stream is sequential -> it lives on ONE owner rank; the only data-path exchange is the
    gather of encoded frame tokens ([n, 49, 3584] bf16, 351,232 B per frame) to that rank (`FrameParallelEncoder`),
    sent batch by batch so the owner can start decoding the first frames while later ones are still being encoded.

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


def videos_for_rank(n_videos, world, rank):
    return list(range(rank, n_videos, world))


class FrameParallelEncoder:
    """encode_fn(frames[b0:b1]) -> tokens [(b1-b0)*tokens_per_frame, hidden].  Every rank encodes its frame range in
    batches; non-owner ranks isend each finished batch, the owner irecvs straight into the frame-ordered output."""

    def __init__(self, encode_fn, tokens_per_frame, hidden, dtype=torch.bfloat16, device="cuda", owner=0, batch=32, group=None):
        self.encode_fn, self.tpf, self.hidden, self.dtype, self.device = encode_fn, tokens_per_frame, hidden, dtype, device
        self.owner, self.batch, self.group = owner, batch, group

    def encode(self, n_frames, local_frames):
        """local_frames: this rank's slice (frame_range) of the video.  Returns (tokens, ready) on the owner — `ready[i]`
        is a callable that blocks until frame i's tokens have arrived — and (None, None) elsewhere."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = frame_range(n_frames, world, rank)
        assert len(local_frames) == hi - lo, (len(local_frames), lo, hi)
        is_owner = rank == self.owner
        out, recvs = None, {}
        if is_owner:
            out = torch.empty(n_frames * self.tpf, self.hidden, dtype=self.dtype, device=self.device)
            for src in range(world):
                if src == rank:
                    continue
                s_lo, s_hi = frame_range(n_frames, world, src)
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


def gather_results(obj, dst=0, group=None):
    """Per-rank Python results (score traces of the rank's videos) -> list on `dst` (a few KB; config 4's only exchange)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(obj, bucket, dst=dst, group=group)
    return bucket
