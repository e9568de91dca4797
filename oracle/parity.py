"""TEST INFRASTRUCTURE ONLY — the oracle as a CHECKER of whole streams (tests/ and bench.py's parity block, outside any
timed region).  Runs oracle/restate.py (fp32 on bf16-rounded weights and pixel inputs: the oracle of record)
over a whole video stream on whatever device the weights live on and reduces a CUDA-path result against it to the four
numbers the north star names: frame-embedding max-abs, score max-abs, identical threshold-crossing frames, and the minimum
|score - threshold| margin that says how much room the crossing comparison had."""
import numpy as np
import torch

from . import restate as R


@torch.no_grad()
def oracle_stream(w, arch, pixels, prefix_ids, frames_per_pass=40, enc_batch=8, query=None):
    """The reference's stream (test/inference.py:196-246) through the restatement: visual_embed in batches, then the decoder
    over [prefix | frame 0 .. frame T-1] with the two heads read at every frame's last token.

    A k-frame causal pass is arithmetically a k-fold single-frame step (tests/test_oracle.py::test_chunked_frames_equal_stepwise,
    6e-8 in fp32), so the oracle may run `frames_per_pass` frames per call to finish in seconds at the full architecture.
    query: optional (frame_index, token_ids) — a user turn encoded before that frame (test/inference.py:281-282).
    Returns dict(emb [T*n, H], scores [T, 2] (informative, relevance), head_logits [T, 4])."""
    dev = pixels.device
    n = arch.frame_tokens
    T = pixels.shape[0]
    emb = torch.cat([R.visual_embed(w, arch, pixels[b:b + enc_batch]) for b in range(0, T, enc_batch)], 0)
    cache = R.KVCache(arch.layers)
    scores, logits = [], []
    f0 = 0
    while f0 < T:
        nf = min(frames_per_pass, T - f0)
        if query is not None and f0 <= query[0] < f0 + nf:
            if query[0] > f0:
                nf = query[0] - f0          # stop the pass in front of the query turn
            else:
                R.model_forward(w, arch, R.embed_tokens(w, torch.as_tensor(query[1], device=dev)), cache)
                query = None
                continue
        pre = R.embed_tokens(w, torch.as_tensor(prefix_ids, device=dev)) if f0 == 0 and len(cache) == 0 and len(prefix_ids) \
            else torch.zeros(0, arch.hidden, device=dev)
        P = pre.shape[0]
        out = R.model_forward(w, arch, torch.cat([pre.to(emb.dtype), emb[f0 * n:(f0 + nf) * n]]), cache)
        rows = torch.tensor([P + n * (j + 1) - 1 for j in range(nf)], device=dev)
        il, rl = out["informative_logits"][rows], out["relevance_logits"][rows]
        scores.append(torch.stack([il.softmax(-1)[:, 1], rl.softmax(-1)[:, 1]], 1))
        logits.append(torch.cat([il, rl], 1))
        del out
        f0 += nf
    return {"emb": emb, "scores": torch.cat(scores, 0), "head_logits": torch.cat(logits, 0)}


def threshold_at_quantile(ref_scores, q=0.8):
    """The single-frame threshold SURVEY.md §8(d) specifies for configs[1]: the oracle's q-quantile score."""
    return float(np.quantile(np.asarray(ref_scores, dtype=np.float64), q))


def crossings(scores, thr):
    """test/inference.py:300-301: need_response when the selected heads' score is strictly greater than the threshold."""
    return [int(i) for i in np.nonzero(np.asarray(scores, dtype=np.float64) > thr)[0]]


def parity_report(ref, got_scores, got_emb=None, head=0, q=0.8):
    """ref: oracle_stream() result; got_scores [T,2]; got_emb [T*n,H] (optional).  All comparisons in float64 on the host.

    Decisions are compared at TWO thresholds: the oracle's q-quantile itself (np.quantile lands INSIDE a gap
    between two neighbouring scores, at 0.2 of it for 120 frames, so its margin can be arbitrarily small) and the middle of the
    widest gap between neighbouring oracle scores in the [q - 0.05, q + 0.05] quantile band.  `flips_outside_noise` counts the
    frames whose decision differs although their oracle score is further from the threshold than the measured score error —
    the number that must be 0 for any implementation within tolerance."""
    rs = ref["scores"].double().cpu().numpy()
    gs = torch.as_tensor(got_scores).double().cpu().numpy()
    err = float(np.abs(rs - gs).max())
    r, g = rs[:, head], gs[:, head]

    def at(thr):
        rc, gc = crossings(r, thr), crossings(g, thr)
        flips = sorted(set(rc) ^ set(gc))
        return {"threshold": thr, "crossings_match": rc == gc, "n_crossings": len(rc), "flips": flips,
                "flips_outside_noise": [i for i in flips if abs(r[i] - thr) > err], "min_margin": float(np.abs(r - thr).min()),
                "crossings_ref": rc[:32], "crossings_got": gc[:32]}
    main = at(threshold_at_quantile(r, q))
    so = np.sort(r)
    lo, hi = int(np.floor((q - 0.05) * (len(so) - 1))), int(np.ceil((q + 0.05) * (len(so) - 1)))
    j = lo + int(np.argmax(so[lo + 1:hi + 1] - so[lo:hi]))
    gap = at(float((so[j] + so[j + 1]) / 2))
    rep = {"score_maxabs": err, "threshold_rule": f"oracle {int(q * 100)}th-percentile of head {head} (0 = informative)", **main,
           "widest_gap_near_quantile": {k: gap[k] for k in ("threshold", "crossings_match", "n_crossings", "flips", "min_margin")}}
    if got_emb is not None:
        re = ref["emb"].float()
        ge = got_emb.float().to(re.device)
        rep["emb_maxabs"] = float((re - ge).abs().max())
        rep["emb_bf16_floor"] = float((re.bfloat16().float() - re).abs().max())   # rounding the EXACT oracle values to bf16
        rep["emb_ref_absmax"] = float(re.abs().max())
    return rep
