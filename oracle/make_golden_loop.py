"""TEST INFRASTRUCTURE ONLY — generates tests/golden/tiny_loop_*.npz by running the reference's OWN frame-loop class
(test/inference.py LiveInferForBenchmark, imported unmodified from /root/reference through oracle/ref_import.py) over the
reference's own model class on CPU in fp32:  python -m oracle.make_golden_loop   (authoring container only).

Each fixture holds, for one seeded TINY-architecture video + query + flag set: the per-frame scores (debug_data), the
response times, the generated token ids of every response, the ids argmax-ed after the query turn, the final context
length (the rollback semantics of remove_assistant_turns), and the smallest top-2 lm-logit gap behind any token decision
(a CUDA-vs-oracle id comparison is only meaningful when that gap exceeds the kernels' logit noise).
tests/test_oracle.py pins oracle/restate.LiveLoopOracle to these; tests/test_gpu_loop.py pins the CUDA loop."""
import json
import os

import numpy as np
import torch

from . import arch as A
from . import ref_import as RI
from . import restate as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (seed, n_frames, flags, query (time, text) or None, max_new_tokens)
    "tiny_loop_keep_turns": (41, 8, dict(stream_end_prob_threshold=None, score_heads="informative_score", remove_assistant_turns=False,
                                          repetition_penalty=1.15), (1.0, "what is happening now"), 3),
    "tiny_loop_rollback": (42, 8, dict(stream_end_prob_threshold=None, score_heads="informative_score,relevance_score",
                                        remove_assistant_turns=True, repetition_penalty=None), (0.0, "describe the scene"), 4),
    "tiny_loop_score_sum": (43, 9, dict(stream_end_score_sum_threshold=None, score_heads="informative_score", remove_assistant_turns=True,
                                         repetition_penalty=1.3), None, 4),
}


def oracle_loop_for(arch, w, loop, flags, frames, query_ids, max_new):
    """oracle/restate.LiveLoopOracle configured like the reference loop object `loop`."""
    kw = {k: v for k, v in flags.items() if k in ("stream_end_prob_threshold", "stream_end_score_sum_threshold", "remove_assistant_turns",
                                                  "repetition_penalty")}
    lo = R.LiveLoopOracle(w, arch, start_ids=loop._start_ids.view(-1).tolist(), stream_prompt_ids=loop._added_stream_prompt_ids.view(-1).tolist(),
                          stream_generation_ids=loop._added_stream_generation_ids.view(-1).tolist(), eos_token_id=loop.eos_token_id,
                          frame_fps=2, max_new_tokens=max_new, score_heads=flags["score_heads"].split(","), **kw)
    lo.input_video_stream(R.preprocess_frames(frames))
    if query_ids is not None:
        lo.input_query_stream([query_ids])
    return lo


def run_case(name, seed, n_frames, flags, query, max_new, min_gap=0.03, min_margin=0.012):
    """Searches seeds from `seed` upward for a case whose every token decision has a top-2 logit gap >= min_gap and whose
    every score is >= min_margin away from the threshold (so that a bf16 implementation can be held to the SAME ids and
    crossings), then stores what the reference's own loop class produced for it."""
    from mmduet_b200.arguments_live import LiveTestArguments
    from mmduet_b200.tokenization_live import SyntheticTokenizer
    arch = A.TINY
    tok = SyntheticTokenizer(arch.vocab)
    for seed in range(seed, seed + 400, 7):
        w = R.make_weights(arch, seed=seed)
        frames = R.synthetic_frames(n_frames, seed=seed + 1)
        fl = dict(flags)

        def make(f2):
            loop = RI.build_reference_loop(arch, w, tok, LiveTestArguments(frame_fps=2, system_prompt="you watch a video", **f2))
            loop.inplace_output_ids = torch.zeros(1, max_new, dtype=torch.long)
            loop.input_video_stream(RI.as_cpu_tensor(frames))
            if query:
                loop.input_query_stream([{"role": "user", "time": query[0], "content": query[1]}])
            return loop
        # pass 1 (never responds) to learn the score distribution, then place the threshold in the widest gap
        probe_flags = {k: v for k, v in fl.items() if not k.endswith("_threshold")}
        probe = make(dict(probe_flags, stream_end_prob_threshold=10.0))
        probe.inference()
        heads = fl["score_heads"].split(",")
        sc = np.array([sum(d[h] for h in heads) for d in probe.debug_data_list])
        if "stream_end_score_sum_threshold" in fl:
            fl["stream_end_score_sum_threshold"] = float(round(2.6 * sc.mean(), 3))
        else:
            so = np.sort(sc)
            j = int(np.argmax(so[1:-2] - so[:-3]))          # at least two responding frames
            fl["stream_end_prob_threshold"] = float((so[j] + so[j + 1]) / 2)
        loop = make(fl)
        responses = loop.inference()
        tokw = loop.tokenizer
        # the restated loop on the same case: must reproduce the reference loop exactly; provides the decision margins
        q_ids = None
        if query:
            q_ids = (query[0], lambda role: tok.apply_chat_template([{"role": "user", "content": query[1]}],
                                                                    add_stream_query_prompt=role == "stream", add_stream_prompt=True))
        lo = oracle_loop_for(arch, w, loop, fl, frames, q_ids, max_new)
        lo_resp = lo.inference()
        assert [r["content"] for r in lo_resp] == tokw.decoded, (lo_resp, tokw.decoded)
        assert [r["time"] for r in lo_resp] == [r["time"] for r in responses if r["role"] == "assistant"]
        got = np.array([[d["informative_score"], d["relevance_score"]] for d in loop.debug_data_list])
        ours = np.array([[d["informative_score"], d["relevance_score"]] for d in lo.debug_data_list])
        assert np.abs(got - ours).max() < 1e-5
        gap = min([g for _, g in lo.top2_gaps] or [np.inf])
        s2 = np.array([sum(d[h] for h in heads) for d in loop.debug_data_list])
        if "stream_end_prob_threshold" in fl:
            margin = float(np.abs(s2 - fl["stream_end_prob_threshold"]).min())
        else:
            acc, margin = 0.0, 1e9
            for v in s2:
                acc += v
                margin = min(margin, abs(acc - fl["stream_end_score_sum_threshold"]))
                if acc > fl["stream_end_score_sum_threshold"]:
                    acc = 0.0
        n_resp = sum(r["role"] == "assistant" for r in responses)
        print(f"  {name}: seed {seed}: {n_resp} responses, min top-2 gap {gap:.4f}, min score margin {margin:.4f}")
        if gap >= min_gap and margin >= min_margin and 2 <= n_resp < n_frames:
            break
    else:
        raise RuntimeError("no seed found")
    flags = fl
    fx = dict(
        seed=seed, n_frames=n_frames, flags=json.dumps(flags), max_new_tokens=max_new,
        query_time=-1.0 if query is None else query[0], query_text="" if query is None else query[1],
        scores=got.astype(np.float64),
        times=np.array([d["time"] for d in loop.debug_data_list], dtype=np.float64),
        response_times=np.array([r["time"] for r in responses if r["role"] == "assistant"], dtype=np.float64),
        generated=np.array([g + [-1] * (max_new - len(g)) for g in tokw.decoded], dtype=np.int64).reshape(-1, max_new),
        generated_token_ids=np.array(loop.generated_token_ids, dtype=np.int64),
        final_context=int(loop.past_key_values.get_seq_length()),
        last_ids=loop.last_ids.view(-1).numpy().astype(np.int64),
        min_top2_gap=gap, min_score_margin=margin)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)
    print(name, "flags", flags, "responses at", fx["response_times"], "generated", fx["generated"].tolist(),
          "final ctx", fx["final_context"], "min top-2 gap", gap, "score margin", margin)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    torch.set_num_threads(8)
    for name, (seed, n, flags, query, max_new) in CASES.items():
        run_case(name, seed, n, flags, query, max_new)


if __name__ == "__main__":
    main()
