"""Tiny vision scenario for compute-sanitizer (initcheck / memcheck): visual_embed at 1 and 3 frames + legacy entry."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200.config import ModelConfig
from mmduet_b200.engine import VisionEngine
from oracle import arch as A, restate as R

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
arch = A.TINY
w = R.make_weights(arch, seed=11)
cfg = ModelConfig.from_any(arch)
vis = VisionEngine(cfg, w, dev)
frames = R.synthetic_frames(3, seed=12).to(dev)
for T in (1, 3):
    e = vis.visual_embed(frames[:T], normalize=True)
    torch.cuda.synchronize()
    print(T, float(e.float().abs().sum()))
