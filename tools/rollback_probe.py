"""Probe for the memcheck-only mismatch of tests/test_gpu_loop.py::test_loop_matches_oracle_with_rollback:
(a) two identical no-generation runs, (b) two identical generation+rollback runs, (c) generation vs none."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_grad_enabled(False)
from mmduet_b200 import build_model_and_tokenizer
from mmduet_b200.arguments_live import LiveTestArguments
from mmduet_b200.config import ModelConfig
from mmduet_b200.inference import LiveInferForBenchmark
from oracle import arch as A, restate as R
arch = A.SMALL
w = R.make_weights(arch, seed=91)
model, tok = build_model_and_tokenizer(state_dict=w, model_config=ModelConfig.from_any(arch), device="cuda:0", max_context=4096)
frames = R.synthetic_frames(14, seed=5)

def run(thr, **kw):
    inf = LiveInferForBenchmark(LiveTestArguments(frame_fps=2, system_prompt="a b c d e f", stream_end_prob_threshold=thr, **kw), model=model, tokenizer=tok)
    inf.inplace_output_ids = torch.zeros(1, 3, device=inf.device, dtype=torch.long)
    inf.input_video_stream(frames)
    resp = inf.inference()
    return np.array([[d["informative_score"], d["relevance_score"]] for d in inf.debug_data_list]), [r["time"] for r in resp]

a1, _ = run(1.0)
a2, _ = run(1.0)
s = np.sort(a1[:, 0]); j = int(np.argmax(s[1:] - s[:-1])); thr = float((s[j] + s[j + 1]) / 2)
b1, t1 = run(thr, remove_assistant_turns=True)
b2, t2 = run(thr, remove_assistant_turns=True)
a3, _ = run(1.0)
print("none vs none      ", np.abs(a1 - a2).max(), np.abs(a1 - a3).max())
print("gen  vs gen       ", np.abs(b1 - b2).max(), t1 == t2, t1)
print("gen  vs none      ", np.abs(b1 - a1).max(), np.abs(b1 - a1).max(1).round(6).tolist())
