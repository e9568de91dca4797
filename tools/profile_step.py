"""Small, fixed kernel sequence for ncu (launch list / --set full captures).  Never a source of bench numbers.
  python tools/profile_step.py step        one 32-frame encode + two 10-frame decoder passes at ~3k context (full-size model)
  python tools/profile_step.py gate_up     the dominant weight-streaming GEMM alone (M=1960 = 40 frames per pass, and M=49), 4 launches each
  python tools/profile_step.py vit         one ViT layer's kernels at batch 32
  python tools/profile_step.py stream40    bench.py's default stream shape: 40-frame encode + one 40-frame decoder pass (1960 tokens) at ~2k context
  python tools/profile_step.py k1          live mode: one single-frame encode + 3 single-frame decoder steps at ~3k context
Wrap in `ncu --nvtx --nvtx-include "profiled/" ...` to keep only the kernels of the NVTX range."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import ops  # noqa: E402
from mmduet_b200.config import ModelConfig  # noqa: E402
from mmduet_b200.engine import DecoderEngine, VisionEngine  # noqa: E402
from mmduet_b200.random_init import random_state_dict, synthetic_frames  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "step"
dev = torch.device("cuda:0")
torch.manual_seed(0)
if mode == "gate_up":
    H, I = 3584, 18944
    wg = [(torch.randn(I, H, device=dev) * 0.02).bfloat16() for _ in range(3)]
    wu = [(torch.randn(I, H, device=dev) * 0.02).bfloat16() for _ in range(3)]
    wil = [torch.stack([g, u], 1).reshape(2 * I, H).contiguous() for g, u in zip(wg, wu)]   # row 2j = gate_j, 2j+1 = up_j
    for M in (1960, 49):
        x = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
        out = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
        for i in range(4):
            if M > 128:   # what the decoder step launches above 128 tokens: one interleaved operand, 256-token tiles
                ops.gemm_t_swiglu_interleaved(x, wil[i % 3], out=out)
            else:
                ops.gemm_t_swiglu(x, wg[i % 3], wu[i % 3], out=out)
    torch.cuda.synchronize()
    sys.exit(0)

cfg = ModelConfig()
sd = random_state_dict(cfg, seed=1234, device=dev, include_lm_head=False)
if mode == "vit":
    cfg1 = ModelConfig(vit_layers_total=2)
    vis = VisionEngine(cfg1, sd, dev)
    fr = synthetic_frames(32, seed=1, device=dev)
    for _ in range(3):
        vis.visual_embed(fr, normalize=True)
    torch.cuda.synchronize()
    sys.exit(0)

if mode in ("stream40", "k1"):
    vis = VisionEngine(cfg, sd, dev)
    dec = DecoderEngine(cfg, sd, dev, max_context=8192, max_tokens=2048)
    n_f = 40 if mode == "stream40" else 1
    fr = synthetic_frames(40, seed=1, device=dev)
    emb = vis.visual_embed(fr, normalize=True)
    st, L = dec.new_stream(), 0
    rows40 = [49 * (j + 1) - 1 for j in range(40)]
    for p in range(2 if mode == "stream40" else 1):   # warm-up: ~2k (stream40: 4k) tokens of context, every kernel variant seen once
        out = dec.step([dict(storage=st, past=L, ids=[], frames=emb, score_rows=rows40)], score="frame_ends")
        L = out["views"][0].length
    if mode == "k1":
        for f in range(25):
            out = dec.step([dict(storage=st, past=L, ids=[], frames=emb[f * 49:(f + 1) * 49])])
            L = out["views"][0].length
        vis.visual_embed(fr[:1], normalize=True)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("profiled")
    e = vis.visual_embed(fr[:n_f], normalize=True)
    if mode == "stream40":
        out = dec.step([dict(storage=st, past=L, ids=[], frames=e, score_rows=rows40)], score="frame_ends")
    else:
        for f in range(3):
            out = dec.step([dict(storage=st, past=L, ids=[], frames=emb[f * 49:(f + 1) * 49])])
            L = out["views"][0].length
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("context", out["views"][0].length, "scores", out["scores"][:2].tolist())
    sys.exit(0)

vis = VisionEngine(cfg, sd, dev)
dec = DecoderEngine(cfg, sd, dev, max_context=8192, max_tokens=512)
fr = synthetic_frames(32, seed=1, device=dev)
emb = vis.visual_embed(fr, normalize=True)
st = dec.new_stream()
L = 0
g = torch.Generator(device=dev).manual_seed(3)
fill = (torch.randn(490, cfg.hidden, device=dev, generator=g) * 1.1).bfloat16()
for _ in range(6):   # context ~2.9k tokens (untimed warm-up, also warms every kernel)
    out = dec.step([dict(storage=st, past=L, embeds=fill)], score="last")
    L = out["views"][0].length
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("profiled")
emb = vis.visual_embed(fr, normalize=True)
for p in range(2):
    rows = [49 * (j + 1) - 1 for j in range(10)]
    out = dec.step([dict(storage=st, past=L, ids=[], frames=emb[p * 490:(p + 1) * 490], score_rows=rows)], score="frame_ends")
    L = out["views"][0].length
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("context", L, "scores", out["scores"][:2].tolist())
