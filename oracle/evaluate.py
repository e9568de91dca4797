"""TEST INFRASTRUCTURE ONLY — numpy/Python restatement of the score post-processing in the reference's grounding evaluator
(test/evaluate.py): smooth_pred_list (:166-167), normalize_pred_list (:170-173), is_time_in_span (:102-106), calculate_iou
(:129-137) and the window x threshold sweep with its summary lines (:374-399).  Pinned against the reference's own function
bodies by tests/test_postprocess_cpu.py::test_oracle_matches_reference_functions (authoring container) and used as the
checker of mmd_grounding_sweep."""
import numpy as np

WINDOWS = list(range(0, 15))                      # test/evaluate.py:374
THRESHOLDS = np.arange(0.30, 0.71, 0.02)          # test/evaluate.py:376


def smooth_pred_list(pred_list, window_size=4):
    return [np.mean(pred_list[max(0, i - window_size):min(len(pred_list), i + window_size + 1)]) for i in range(len(pred_list))]


def normalize_pred_list(pred_list):
    max_num, min_num = max(pred_list), min(pred_list)
    return [(p - min_num) / (max_num - min_num) for p in pred_list]


def is_time_in_span(time, spans):
    return any(span[0] <= time <= span[1] for span in spans)


def calculate_iou(pred_scores, gold_scores, threshold):
    assert len(pred_scores) == len(gold_scores)
    pred = [p >= threshold for p in pred_scores]
    inter = sum(p and g for p, g in zip(pred, gold_scores))
    union = sum(p or g for p, g in zip(pred, gold_scores))
    return 0 if union == 0 else inter / union


def debug_entry(e):
    """(time, relevance score) of one debug_data entry, in either format the reference writes: the live loop's
    {'time', 'relevance_score': float} (test/inference.py:285) or the deprecated loop's {'video_time',
    'relevance_score': [p0, p1]} that test/evaluate.py:381-386 indexes."""
    t = e["video_time"] if "video_time" in e else e["time"]
    if "relevance_score" not in e:
        return t, 0
    s = e["relevance_score"]
    return t, (s[1] if isinstance(s, (list, tuple)) else s)


def grounding_sweep(pred_examples, gold_examples):
    """test/evaluate.py:374-399.  Returns (final_results, best) with final_results a list of
    {'smooth_window_size', 'threshold', 'scores': [mean IoU, R@0.3, R@0.5, R@0.7]} and best[w] the 'best among all
    thresholds' line of window w."""
    final_results, best = [], {}
    for w in WINDOWS:
        ious = {float(t): [] for t in THRESHOLDS}
        for ex in pred_examples:
            gold = gold_examples[ex["question_id"]]
            times, scores = zip(*[debug_entry(e) for e in ex["debug_data"]])
            with np.errstate(invalid="ignore"):      # a window wider than the video: np.float64 0/0 = nan, as in the evaluator
                p = normalize_pred_list(smooth_pred_list(list(scores), w))
            g = [is_time_in_span(t, gold["timestamps"]) for t in times]
            for t in ious:
                ious[t].append(calculate_iou(p, g, t))
        for t, lst in ious.items():
            final_results.append({"smooth_window_size": w, "threshold": t,
                                  "scores": [np.mean(lst) * 100] + [np.mean([e >= r for e in lst]) * 100 for r in (0.3, 0.5, 0.7)]})
        top = [max(lst[i] for lst in ious.values()) for i in range(len(pred_examples))]
        best[w] = [np.mean(top) * 100] + [np.mean([e >= r for e in top]) * 100 for r in (0.3, 0.5, 0.7)]
    return final_results, best
