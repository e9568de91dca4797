"""Live-mode decoder step latency in isolation: one single-frame (49-token) KV-append step + heads per call, host wall clock with
a synchronize per step (what a live caller sees), over a 120-frame stream (context 32 -> 5.9k), full architecture, random init.

    python tools/bench_k1_step.py [--frames 120] [--streams 3]      # env switches (MMD_L2_PREFETCH=0, ...) apply"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--streams", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.engine import DecoderEngine
    from mmduet_b200.random_init import random_state_dict
    torch.set_grad_enabled(False)
    dev = torch.device("cuda:0")
    cfg = ModelConfig()
    sd = {k: v for k, v in random_state_dict(cfg, seed=1234, device=dev, include_lm_head=False).items() if "vision_tower" not in k}
    dec = DecoderEngine(cfg, sd, dev, max_context=8192, max_tokens=256)
    del sd
    g = torch.Generator(device=dev).manual_seed(3)
    fe = (torch.randn(a.frames * 49, cfg.hidden, generator=g, device=dev) * 1.14).bfloat16()
    prefix = list(range(100, 132))
    ms = []
    for s in range(a.streams + 1):                      # stream 0 is the warm-up
        st, L = dec.new_stream(), 0
        for f in range(a.frames):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            o = dec.step([dict(storage=st, past=L, ids=prefix if f == 0 else [], frames=fe[f * 49:(f + 1) * 49])])
            sc = o["scores"].cpu()
            t1 = time.perf_counter()
            L = o["views"][0].length
            if s > 0 and f > 0:
                ms.append((t1 - t0) * 1e3)
        st.release()
    ms = np.array(ms)
    r = {"p50_ms": float(np.percentile(ms, 50)), "p10_ms": float(np.percentile(ms, 10)), "p99_ms": float(np.percentile(ms, 99)),
         "mean_ms": float(ms.mean()), "n": int(ms.size), "final_context": L,
         "env": {k: v for k, v in os.environ.items() if k.startswith("MMD_")}}
    print(json.dumps(r))
    if a.out:
        json.dump(r, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
