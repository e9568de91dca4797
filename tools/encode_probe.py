"""Is the first visual_embed of a process equal to the second (memcheck showed a first-run-only difference)?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_grad_enabled(False)
from mmduet_b200.config import ModelConfig
from mmduet_b200.engine import VisionEngine
from oracle import arch as A, restate as R
arch = A.SMALL
w = R.make_weights(arch, seed=91)
cfg = ModelConfig.from_any(arch)
vis = VisionEngine(cfg, w, torch.device("cuda:0"))
from mmduet_b200 import _lib
if os.environ.get("ATTN_IMPL"):
    _lib.load().mmd_set_attention_impl(int(os.environ["ATTN_IMPL"]))
if os.environ.get("GEMM_2CTA"):
    _lib.load().mmd_set_gemm_2cta(int(os.environ["GEMM_2CTA"]))
frames = R.synthetic_frames(14, seed=5).cuda()
print("vit heads", cfg.vit_heads, "dim", cfg.vit_dim, "layers", cfg.vit_layers_total)
r1 = vis.tower(frames, True).clone(); e1 = vis.visual_embed(frames, normalize=True).clone()
r2 = vis.tower(frames, True).clone(); e2 = vis.visual_embed(frames, normalize=True).clone()
r3 = vis.tower(frames, True).clone()
S = cfg.patches
d = (r1 - r2).abs().view(14, S, -1)
print("tower first vs second: max", float(d.max()), "per frame", [round(float(x), 6) for x in d.amax((1, 2))])
print("tower second vs third:", float((r2 - r3).abs().max()), " embed first vs second:", float((e1.float() - e2.float()).abs().max()))
if float(d.max()) > 0:
    fr = int(d.amax((1, 2)).argmax()); rows = (d[fr].amax(1) > 0).nonzero().flatten()
    print("frame", fr, "differing rows", rows[:10].tolist(), "...", int(rows.numel()), "of", S)
n = int(os.environ.get("STRESS", "0"))
if n:
    bad = 0
    for i in range(n):
        r = vis.tower(frames, True)
        if not torch.equal(r, r2):
            bad += 1
    torch.cuda.synchronize()
    print("stress", n, "tower runs, mismatches vs run 2:", bad)
