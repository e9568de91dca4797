"""torchrun --nproc-per-node N tools/check_multigpu.py : frame-parallel encoder + NCCL gather to the decoder owner,
checked against a single-rank encode of the same frames (BASELINE configs[2] shape, shortened)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200.config import ModelConfig  # noqa: E402
from mmduet_b200.engine import DecoderEngine, VisionEngine  # noqa: E402
from mmduet_b200.parallel import FrameParallelEncoder, PeerStoreEncoder, frame_range  # noqa: E402
from mmduet_b200.random_init import random_state_dict, synthetic_frames  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = ModelConfig()
sd = random_state_dict(cfg, seed=1234, device=dev, include_lm_head=False)
vis = VisionEngine(cfg, sd, dev)
n = int(os.environ.get("N_FRAMES", "96"))
frames = synthetic_frames(n, seed=5, device=dev)          # same seed on every rank: the full video, sliced below
lo, hi = frame_range(n, world, rank)
mode = os.environ.get("MMD_EXCHANGE", "peer")              # "peer": kernel stores into the owner's HBM; "nccl": send/recv
if mode == "peer":
    enc = PeerStoreEncoder(lambda fr, dst: vis.visual_embed(fr, normalize=True, out=dst), vis.tokens_per_frame, cfg.hidden,
                           max_frames=n, device=dev, owner=0)
else:
    enc = FrameParallelEncoder(lambda fr: vis.visual_embed(fr, normalize=True), vis.tokens_per_frame, cfg.hidden, device=dev, owner=0)
for _ in range(2):
    out, ready = enc.encode(n, frames[lo:hi])
    FrameParallelEncoder.wait_all(ready)
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
out, ready = enc.encode(n, frames[lo:hi])
FrameParallelEncoder.wait_all(ready)
torch.cuda.synchronize()
dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    ref = vis.visual_embed(frames, normalize=True)
    diff = (out.float() - ref.float()).abs().max().item()
    print(f"world {world} [{mode}]: {n} frames encoded+gathered in {dt*1e3:.1f} ms ({n/dt:.0f} frames/s), max diff vs single-rank encode {diff}")
    assert diff == 0.0
    dec = DecoderEngine(cfg, sd, dev, max_context=n * 49 + 64, max_tokens=512)
    st, L = dec.new_stream(), 0
    for f0 in range(0, n, 8):
        o = dec.step([dict(storage=st, past=L, ids=[], frames=out[f0 * 49:(f0 + 8) * 49], score_rows=[49 * (j + 1) - 1 for j in range(min(8, n - f0))])],
                     score="frame_ends")
        L = o["views"][0].length
    print("owner decoded", L, "tokens; last scores", o["scores"][-1].tolist())
dist.destroy_process_group()
