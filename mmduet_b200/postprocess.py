"""Output side of the path: the grounding evaluator's score post-processing — smoothing windows 0..14,
min-max normalisation, threshold sweep 0.30..0.70 (test/evaluate.py:166-173, 363-399) — batched over all videos, windows and
thresholds in ONE kernel launch (csrc/postprocess.cu, float64, numpy's summation order: bit-identical to the evaluator).
Consumes the JSONL records run_benchmark.py writes (debug_data rounded to 3 decimals) and returns the evaluator's result
structures; the few means over videos at the end are the evaluator's own numpy expressions on the exact integer counts."""
import numpy as np
import torch

from . import _lib

WINDOWS = list(range(0, 15))
THRESHOLDS = np.arange(0.30, 0.71, 0.02)


def _entry(e):
    t = e["video_time"] if "video_time" in e else e["time"]
    if "relevance_score" not in e:
        return t, 0.0
    s = e["relevance_score"]
    return t, float(s[1] if isinstance(s, (list, tuple)) else s)


def sweep_counts(score_lists, gold_lists, windows=WINDOWS, thresholds=THRESHOLDS, device="cuda", return_normalized=False):
    """score_lists / gold_lists: per video, the relevance scores and the 0/1 gold labels of its frames.
    Returns counts int32 [n_windows, n_videos, n_thresholds, 2] (intersection, union) on the host
    (+ the normalised scores [n_windows, n_videos, t_max] float64 when asked).  A constant smoothed list normalises to nan
    (np.float64 0/0, as in the evaluator) and predicts nothing; an empty list raises ValueError like the evaluator's max([])."""
    n_v = len(score_lists)
    t_max = max((len(s) for s in score_lists), default=0)
    if n_v == 0 or t_max == 0:
        raise ValueError("max() arg is an empty sequence")
    sc = np.zeros((n_v, t_max), dtype=np.float64)
    gd = np.zeros((n_v, t_max), dtype=np.uint8)
    lens = np.zeros(n_v, dtype=np.int32)
    for v, (s, g) in enumerate(zip(score_lists, gold_lists)):
        assert len(s) == len(g)
        lens[v] = len(s)
        sc[v, :len(s)] = np.asarray(s, dtype=np.float64)
        gd[v, :len(g)] = np.asarray(g, dtype=bool)
    dev = torch.device(device)
    d_sc, d_gd, d_len = (torch.from_numpy(a).to(dev) for a in (sc, gd, lens))
    d_win = torch.tensor(list(windows), dtype=torch.int32, device=dev)
    d_thr = torch.from_numpy(np.asarray(thresholds, dtype=np.float64)).to(dev)
    n_w, n_t = len(d_win), len(d_thr)
    counts = torch.zeros(n_w, n_v, n_t, 2, dtype=torch.int32, device=dev)
    bad = torch.zeros(n_w, n_v, dtype=torch.int32, device=dev)
    norm = torch.zeros(n_w, n_v, t_max, dtype=torch.float64, device=dev) if return_normalized else None
    lib = _lib.load()
    _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())   # fails loudly off a B200
    rc = lib.mmd_grounding_sweep(d_sc.data_ptr(), d_gd.data_ptr(), d_len.data_ptr(), n_v, t_max, d_win.data_ptr(), n_w, d_thr.data_ptr(),
                                 n_t, counts.data_ptr(), _lib.ptr(norm), bad.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, "mmd_grounding_sweep")
    if bool(bad.any()):
        raise ValueError("max() arg is an empty sequence")
    return (counts.cpu().numpy(), norm.cpu().numpy()) if return_normalized else counts.cpu().numpy()


def grounding_sweep(pred_examples, gold_examples, device="cuda"):
    """test/evaluate.py:363-399 for online models: pred_examples = the JSONL records (question_id, debug_data), gold_examples =
    {question_id: {'timestamps': [[start, end], ...]}}.  Returns (final_results, best) exactly as the evaluator builds them."""
    scores, golds = [], []
    for ex in pred_examples:
        spans = gold_examples[ex["question_id"]]["timestamps"]
        ts = [_entry(e) for e in ex["debug_data"]]
        scores.append([s for _, s in ts])
        golds.append([any(a <= t <= b for a, b in spans) for t, _ in ts])
    counts = sweep_counts(scores, golds, device=device).astype(np.float64)
    inter, union = counts[..., 0], counts[..., 1]
    with np.errstate(invalid="ignore", divide="ignore"):
        iou = np.where(union == 0, 0.0, inter / np.where(union == 0, 1.0, union))        # [windows, videos, thresholds]
    final_results, best = [], {}
    for wi, w in enumerate(WINDOWS):
        for ti, t in enumerate(THRESHOLDS):
            lst = iou[wi, :, ti].tolist()
            final_results.append({"smooth_window_size": w, "threshold": float(t),
                                  "scores": [np.mean(lst) * 100] + [np.mean([e >= r for e in lst]) * 100 for r in (0.3, 0.5, 0.7)]})
        top = iou[wi].max(axis=1).tolist()
        best[w] = [np.mean(top) * 100] + [np.mean([e >= r for e in top]) * 100 for r in (0.3, 0.5, 0.7)]
    return final_results, best
