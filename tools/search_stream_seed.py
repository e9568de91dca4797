"""Conditioning of the decision test on the bench stream: for candidate seeds of the synthetic 120-frame stream, run the ORACLE
(fp32 restatement, nothing of the CUDA path) and report how far its 80th-percentile informative threshold sits from the nearest
oracle score.  np.quantile lands 0.2 into a gap between two neighbouring scores, so for most streams that margin is far below
any bf16 implementation's score error and the "identical crossing frames" comparison is a coin toss on one frame; bench.py and
the tests use a stream whose margin exceeds the tolerance-scale error.  The choice depends on the oracle's numbers only.

    python tools/search_stream_seed.py --seeds 1-60 --out gpurun_out/stream_seeds.json"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="1-40")
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    lo, hi = (int(x) for x in a.seeds.split("-"))
    from oracle import arch as A, parity as P, restate as R
    from mmduet_b200.random_init import synthetic_frames
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_grad_enabled(False)
    dev = torch.device("cuda:0")
    w = R.make_weights(A.FULL, seed=1234, device=dev, generate_on_device=True, include_lm_head=False)
    prefix = list(range(100, 132))
    rows = []
    for seed in range(lo, hi + 1):
        px = R.preprocess_frames(synthetic_frames(a.frames, seed=seed, device=dev)).bfloat16().float()
        ref = P.oracle_stream(w, A.FULL, px, prefix, frames_per_pass=40)
        r = ref["scores"][:, 0].double().cpu().numpy()
        thr = P.threshold_at_quantile(r)
        so = np.sort(r)
        row = {"seed": seed, "threshold": thr, "min_margin": float(np.abs(r - thr).min()), "n_crossings": int((r > thr).sum()),
               "score_min": float(so[0]), "score_max": float(so[-1])}
        print(row, flush=True)
        rows.append(row)
        del ref, px
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
