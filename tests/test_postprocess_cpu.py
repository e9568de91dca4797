"""CPU tests of the evaluator post-processing oracle (oracle/evaluate.py): against the reference's own function bodies where
/root/reference exists, and the numpy summation-order model the CUDA kernel implements against np.mean itself."""
import ast
import os

import numpy as np
import pytest

from oracle import evaluate as E
from oracle.ref_import import REFERENCE_ROOT


def _records(rng, n_videos, legacy=False):
    preds, golds = [], {}
    for v in range(n_videos):
        n = int(rng.integers(2, 90))
        fps = float(rng.choice([1.0, 2.0]))
        dbg = []
        for i in range(n):
            s = round(float(rng.random()), 3)
            dbg.append({"video_time": i / fps, "relevance_score": [round(1 - s, 3), s]} if legacy else
                       {"time": i / fps, "informative_score": round(float(rng.random()), 3), "relevance_score": s})
        a = float(rng.random()) * n / fps * 0.7
        preds.append({"question_id": f"q{v}", "debug_data": dbg})
        golds[f"q{v}"] = {"timestamps": [[round(a, 2), round(a + float(rng.random()) * n / fps * 0.3 + 0.5, 2)]]}
    return preds, golds


def numpy_order_mean(a):
    """The summation order csrc/postprocess.cu implements (numpy's pairwise_sum for n <= 128, then / n)."""
    n = len(a)
    if n < 8:
        r = 0.0
        for x in a:
            r += x
        return r / n
    r = list(a[:8])
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] += a[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res += a[i]
        i += 1
    return res / n


def test_summation_order_model_is_numpys():
    rng = np.random.default_rng(0)
    for _ in range(5000):
        a = [round(float(x), 3) for x in rng.random(int(rng.integers(1, 30)))]
        assert numpy_order_mean(a) == float(np.mean(a))


def test_sweep_shapes_and_edge_cases():
    rng = np.random.default_rng(1)
    preds, golds = _records(rng, 5)
    final, best = E.grounding_sweep(preds, golds)
    assert len(final) == 15 * 21 and set(best) == set(range(15))
    assert all(0 <= r["scores"][0] <= 100 for r in final)
    assert abs(E.THRESHOLDS[0] - 0.30) < 1e-12 and abs(E.THRESHOLDS[-1] - 0.70) < 1e-9 and len(E.THRESHOLDS) == 21
    with pytest.raises(ZeroDivisionError):
        E.normalize_pred_list([0.5, 0.5, 0.5])
    assert E.smooth_pred_list([1.0, 2.0, 4.0], 0) == [1.0, 2.0, 4.0]
    assert E.smooth_pred_list([1.0, 2.0, 4.0], 1) == [1.5, 7.0 / 3.0, 3.0]
    assert E.calculate_iou([0.1, 0.2], [False, False], 0.5) == 0


@pytest.mark.reference
def test_oracle_matches_reference_functions():
    """The reference's own function bodies (parsed out of /root/reference/test/evaluate.py: the module itself imports LLM
    judges and Java-backed metrics at import time) against the restatement, on both debug_data formats."""
    src = open(os.path.join(REFERENCE_ROOT, "test", "evaluate.py")).read()
    want = {"smooth_pred_list", "normalize_pred_list", "is_time_in_span", "calculate_iou", "keep_longest_true_span"}
    mod = ast.Module(body=[n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in want], type_ignores=[])
    ns = {"np": np}
    exec(compile(mod, "reference:test/evaluate.py", "exec"), ns)
    rng = np.random.default_rng(2)
    preds, golds = _records(rng, 6, legacy=True)
    for ex in preds:
        times = [e["video_time"] for e in ex["debug_data"]]
        scores = [e["relevance_score"][1] for e in ex["debug_data"]]
        assert [E.debug_entry(e) for e in ex["debug_data"]] == list(zip(times, scores))
        for w in (0, 1, 4, 14):
            a, b = ns["smooth_pred_list"](scores, w), E.smooth_pred_list(scores, w)
            assert a == b
            with np.errstate(invalid="ignore"):      # a window wider than the video: constant list, 0/0 = nan on both sides
                assert np.array_equal(ns["normalize_pred_list"](a), E.normalize_pred_list(b), equal_nan=True)
            g = [ns["is_time_in_span"](t, golds[ex["question_id"]]["timestamps"]) for t in times]
            assert g == [E.is_time_in_span(t, golds[ex["question_id"]]["timestamps"]) for t in times]
            with np.errstate(invalid="ignore"):
                p = E.normalize_pred_list(b)
            for t in E.THRESHOLDS:
                assert ns["calculate_iou"](p, g, t) == E.calculate_iou(p, g, t)
