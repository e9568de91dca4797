"""Encoder latency in isolation: T uint8 frames resident on the GPU -> SigLIP + projector + pooling -> frame tokens, CUDA events,
full architecture, random init.    python tools/bench_encode.py --frames 1 2 4 8 16 40"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="+", default=[1, 2, 4, 8, 16, 40])
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.engine import VisionEngine
    from mmduet_b200.random_init import random_state_dict, synthetic_frames
    torch.set_grad_enabled(False)
    dev = torch.device("cuda:0")
    cfg = ModelConfig()
    sd = {k: v for k, v in random_state_dict(cfg, seed=1234, device=dev, include_lm_head=False).items() if not k.startswith("model.layers")}
    vis = VisionEngine(cfg, sd, dev)
    res = []
    for T in a.frames:
        fr = synthetic_frames(T, seed=1, device=dev)
        for _ in range(3):
            vis.visual_embed(fr, normalize=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            vis.visual_embed(fr, normalize=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        res.append({"frames": T, "ms": round(ms, 4), "frames_per_s": round(T / ms * 1e3, 1)})
        print(res[-1], flush=True)
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("MMD_")}, "results": res}))


if __name__ == "__main__":
    main()
