"""GPU tests of the frame loop (host mirror of test/inference.py / demo/liveinfer.py over the CUDA model) against the
oracle's restatement of the same loop: per-frame scores, threshold-crossing frames, rollback, multi-frame passes."""
import numpy as np
import pytest
import torch

from oracle import arch as A
from oracle import restate as R

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.set_grad_enabled(False)
    from mmduet_b200 import build_model_and_tokenizer
    from mmduet_b200.config import ModelConfig
    arch = A.SMALL
    w = R.make_weights(arch, seed=91)
    # a roomy KV pool: the LiveInfer objects of earlier tests give their pages back only when they are collected
    model, tok = build_model_and_tokenizer(state_dict=w, model_config=ModelConfig.from_any(arch), device="cuda:0", max_context=4096,
                                           kv_pages=1024)
    frames = R.synthetic_frames(14, seed=5)
    return arch, w, model, tok, frames


def _args(**kw):
    from mmduet_b200.arguments_live import LiveTestArguments
    return LiveTestArguments(frame_fps=2, system_prompt="a b c d e f", **kw)


def _oracle_loop(arch, w, infer, frames, **kw):
    loop = R.LiveLoopOracle(w, arch, start_ids=infer._start_ids.view(-1).tolist(),
                            stream_prompt_ids=infer._added_stream_prompt_ids.view(-1).tolist(),
                            stream_generation_ids=infer._added_stream_generation_ids.view(-1).tolist(),
                            eos_token_id=infer.eos_token_id, frame_fps=2, max_new_tokens=infer.inplace_output_ids.shape[1], **kw)
    loop.input_video_stream(R.preprocess_frames(frames).bfloat16().float())
    return loop


def test_loop_matches_oracle_with_rollback(setup):
    from mmduet_b200.inference import LiveInferForBenchmark
    arch, w, model, tok, frames = setup
    # 1) grounding-style probe: threshold 1 never generates (charades.sh:11); collect the oracle's scores
    infer = LiveInferForBenchmark(_args(stream_end_prob_threshold=1.0), model=model, tokenizer=tok)
    infer.inplace_output_ids = torch.zeros(1, 3, device=infer.device, dtype=torch.long)
    infer.input_video_stream(frames)
    assert infer.inference() == []
    probe = _oracle_loop(arch, w, infer, frames, stream_end_prob_threshold=1.0)
    assert probe.inference() == []
    ref = np.array([[d["informative_score"], d["relevance_score"]] for d in probe.debug_data_list])
    got = np.array([[d["informative_score"], d["relevance_score"]] for d in infer.debug_data_list])
    assert [d["time"] for d in infer.debug_data_list] == [d["time"] for d in probe.debug_data_list]
    err = np.abs(ref - got).max()
    assert err < TOL, err
    # 2) a threshold in the widest score gap: identical crossing frames, generation with rollback keeps the scores
    s = np.sort(ref[:, 0])
    j = int(np.argmax(s[1:] - s[:-1]))
    thr = float((s[j] + s[j + 1]) / 2)
    assert (s[j + 1] - s[j]) / 2 > err
    infer2 = LiveInferForBenchmark(_args(stream_end_prob_threshold=thr, remove_assistant_turns=True), model=model, tokenizer=tok)
    infer2.inplace_output_ids = torch.zeros(1, 3, device=infer2.device, dtype=torch.long)
    infer2.input_video_stream(frames)
    resp = infer2.inference()
    loop2 = _oracle_loop(arch, w, infer2, frames, stream_end_prob_threshold=thr, remove_assistant_turns=True)
    resp_ref = loop2.inference()
    assert [r["time"] for r in resp] == [r["time"] for r in resp_ref]
    assert len(resp) == int((ref[:, 0] > thr).sum()) > 0
    got2 = np.array([[d["informative_score"], d["relevance_score"]] for d in infer2.debug_data_list])
    assert np.abs(got2 - got).max() < 1e-6          # rollback: the context is untouched by the generated turns
    # 3) running-sum mode (youcook2.sh:14) on both heads
    kw = dict(stream_end_score_sum_threshold=1.3, score_heads="informative_score,relevance_score", remove_assistant_turns=True)
    infer3 = LiveInferForBenchmark(_args(**kw), model=model, tokenizer=tok)
    infer3.inplace_output_ids = torch.zeros(1, 2, device=infer3.device, dtype=torch.long)
    infer3.input_video_stream(frames)
    resp3 = infer3.inference()
    # same rule applied to the oracle's scores (crossings may only differ if a running sum lands within err of the threshold)
    acc, want, margin = 0.0, [], 1.0
    for i, (a, b) in enumerate(ref):
        acc += a + b
        margin = min(margin, abs(acc - 1.3))
        if acc > 1.3:
            want.append(i / 2.0)
            acc = 0.0
    if margin > 20 * err:
        assert [r["time"] for r in resp3] == want
    with pytest.raises(ValueError):
        LiveInferForBenchmark(_args(), model=model, tokenizer=tok)


@pytest.mark.parametrize("keep_turns", [False, True])
def test_multi_frame_passes_equal_single_frame_steps(setup, keep_turns):
    from mmduet_b200.inference import LiveInferForBenchmark
    arch, w, model, tok, frames = setup
    runs = {}
    for k in (1, 4, 5):
        infer = LiveInferForBenchmark(_args(stream_end_prob_threshold=0.5, remove_assistant_turns=not keep_turns), model=model, tokenizer=tok)
        infer.frames_per_step = k
        infer.inplace_output_ids = torch.zeros(1, 3, device=infer.device, dtype=torch.long)
        infer.input_video_stream(frames)
        infer.input_query_stream([{"role": "user", "time": 1.6, "content": "what is happening now"}])
        resp = infer.inference()
        runs[k] = (resp, [(d["time"], d["informative_score"], d["relevance_score"]) for d in infer.debug_data_list],
                   infer.past_key_values.length)
    base = runs[1]
    assert len(base[1]) == len(frames)
    for k in (4, 5):
        resp, dbg, L = runs[k]
        assert [r["time"] for r in resp] == [r["time"] for r in base[0]]
        assert [r["content"] for r in resp] == [r["content"] for r in base[0]]
        assert L == base[2]
        assert np.abs(np.array(dbg) - np.array(base[1])).max() < 5e-3


def test_demo_one_frame_and_query(setup):
    from mmduet_b200.inference import LiveInferForDemo
    arch, w, model, tok, frames = setup
    demo = LiveInferForDemo(_args(stream_end_prob_threshold=1.0), model=model, tokenizer=tok)
    demo.input_video_stream(frames[:4])
    r0 = demo.input_one_frame()
    assert set(r0) == {"frame_idx", "time", "informative_score", "relevance_score", "response"} and r0["frame_idx"] == 1
    L0 = demo.past_key_values.length
    demo.encode_given_query("describe the scene")
    assert demo.past_key_values.length > L0 and demo.last_role == "user" and demo.last_ids.shape == (1, 1)
    r1 = demo.input_one_frame()
    assert r1["frame_idx"] == 2 and r1["response"] is None and r1["time"] == 0.5


def test_forward_surface_matches_reference_contract(setup):
    arch, w, model, tok, frames = setup
    emb = model.visual_embed(frames[:1].cuda())
    assert emb.shape == (49, arch.hidden) and emb.dtype == torch.bfloat16
    out = model(inputs_embeds=emb[None], use_cache=True, past_key_values=None, return_dict=True)
    assert out.logits.shape == (1, 49, arch.vocab) and out.logits.dtype == torch.float32
    assert out.informative_logits.shape == (1, 49, 2) and out.relevance_logits.shape == (1, 49, 2)
    assert out.past_key_values.get_seq_length() == 49
    wd = {k: v.cuda() for k, v in w.items()}
    o = R.model_forward(wd, arch, emb.float(), R.KVCache(arch.layers), want_lm_logits=True)
    assert (out.informative_logits[0] - o["informative_logits"]).abs().max() < 5e-2
    assert (out.logits[0] - o["logits"]).abs().max() < 6e-2
    ids = torch.tensor([[7, 8, 9]], device="cuda")
    e = model.joint_embed(ids, None)
    assert e.shape == (1, 3, arch.hidden)
    out2 = model(inputs_embeds=e, past_key_values=out.past_key_values, use_cache=True, return_dict=True, logits_to_keep="last")
    assert out2.past_key_values.get_seq_length() == 52 and out2.logits.shape == (1, 1, arch.vocab)


def test_benchmark_driver_writes_reference_jsonl(setup, tmp_path):
    """test/inference.py:332-361 schema: one JSON line per video, debug_data rounded to 3 decimals."""
    import json

    from mmduet_b200.inference import LiveInferForBenchmark
    from mmduet_b200.run_benchmark import run
    arch, w, model, tok, frames = setup
    infer = LiveInferForBenchmark(_args(stream_end_prob_threshold=1.0), model=model, tokenizer=tok)
    items = [("vid0", frames[:5], [{"role": "user", "time": 0.0, "content": "what happens"}], 2, 2.5),
             (None, None, None, None, None),
             ("vid1", frames[5:9], [], 1, 4.0)]
    out = tmp_path / "out.jsonl"
    assert run(infer, items, str(out), grounding_mode=True) == 2
    lines = [json.loads(l) for l in open(out)]
    assert [l["question_id"] for l in lines] == ["vid0", "vid1"]
    assert set(lines[0]) == {"question_id", "model_response_list", "video_duration", "debug_data"}
    assert lines[0]["model_response_list"] == [{"time": 0.0, "content": "what happens", "role": "user"}]
    assert len(lines[0]["debug_data"]) == 5 and len(lines[1]["debug_data"]) == 4
    assert [d["time"] for d in lines[1]["debug_data"]] == [0.0, 1.0, 2.0, 3.0]        # fps 1
    for d in lines[0]["debug_data"]:
        assert set(d) == {"time", "informative_score", "relevance_score"}
        assert round(d["informative_score"], 3) == d["informative_score"]


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (symmetric memory over NVLink)")
def test_peer_store_encoder_two_gpus():
    """Frame-parallel encode where each rank's projector/pool kernel stores into the owner's HBM (no collective):
    bit-identical to a single-rank encode; the tool asserts that itself."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MMD_EXCHANGE="peer", N_FRAMES="40")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "check_multigpu.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "max diff vs single-rank encode 0.0" in r.stdout


class _RecordingTokenizer:
    """SyntheticTokenizer that keeps the raw ids of every decoded response."""

    def __init__(self, vocab):
        from mmduet_b200.tokenization_live import SyntheticTokenizer
        self._t, self.decoded = SyntheticTokenizer(vocab), []

    def __getattr__(self, name):
        return getattr(self._t, name)

    def decode(self, ids, **kw):
        self.decoded.append([int(i) for i in ids.tolist()])
        return self._t.decode(ids, **kw)


@pytest.mark.parametrize("k", [1, 3])
@pytest.mark.parametrize("name", ["tiny_loop_keep_turns", "tiny_loop_rollback", "tiny_loop_score_sum"])
def test_cuda_loop_matches_reference_loop_fixture(name, k):
    """The CUDA loop (LiveInferForBenchmark over the kernels) against what the reference's OWN LiveInferForBenchmark produced
    on the same seeded video / query / flags (tests/golden/tiny_loop_*.npz, oracle/make_golden_loop.py): scores within
    2e-2, identical response frames, identical generated token ids (greedy, with and without the HF repetition penalty),
    identical ids carried after the last turn, identical final context length (rollback).  The fixtures were chosen with a
    top-2 logit gap >= 0.03 and a score-to-threshold margin >= 0.012 so that identity is a fair demand of a bf16 path."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.set_grad_enabled(False)
    from mmduet_b200 import build_model_and_tokenizer
    from mmduet_b200.arguments_live import LiveTestArguments
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.inference import LiveInferForBenchmark
    from tests.test_oracle import loop_case
    g, flags, arch, w, frames, _, query = loop_case(name)
    tok = _RecordingTokenizer(arch.vocab)
    model, _ = build_model_and_tokenizer(state_dict=w, model_config=ModelConfig.from_any(arch), device="cuda:0", max_context=2048,
                                         tokenizer=tok)
    infer = LiveInferForBenchmark(LiveTestArguments(frame_fps=2, system_prompt="you watch a video", **flags), model=model, tokenizer=tok)
    infer.inplace_output_ids = torch.zeros(1, int(g["max_new_tokens"]), device=infer.device, dtype=torch.long)
    infer.frames_per_step = k
    infer.input_video_stream(frames)
    if query:
        infer.input_query_stream([{"role": "user", "time": query[0], "content": query[1]}])
    resp = infer.inference()
    got = np.array([[d["informative_score"], d["relevance_score"]] for d in infer.debug_data_list])
    err = np.abs(got - g["scores"]).max()
    print(name, "k", k, "score max-abs", err, "fixture margins: top-2 gap", float(g["min_top2_gap"]), "score", float(g["min_score_margin"]))
    assert err < TOL, err
    assert [d["time"] for d in infer.debug_data_list] == g["times"].tolist()
    assert [r["time"] for r in resp if r["role"] == "assistant"] == g["response_times"].tolist()
    assert tok.decoded == [[t for t in row if t >= 0] for row in g["generated"].tolist()]
    assert infer.past_key_values.length == int(g["final_context"])
    assert infer.last_ids.view(-1).tolist() == g["last_ids"].tolist()
    assert infer.generated_token_ids == g["generated_token_ids"].tolist()
    if query:
        assert [r for r in resp if r["role"] == "user"] == [{"time": query[0], "content": query[1], "role": "user"}]


def test_query_turn_matches_oracle(setup):
    """_encode_query (test/inference.py:248-255): the same chat-templated user turn goes into the oracle loop and into the
    CUDA loop after two frames; the token kept from the turn (argmax of the last position's lm logits) and the scores of
    the following frames must agree.  The oracle's top-2 logit gap is reported next to the lm-logit error."""
    from mmduet_b200.inference import LiveInferForBenchmark
    arch, w, model, tok, frames = setup
    infer = LiveInferForBenchmark(_args(stream_end_prob_threshold=1.0), model=model, tokenizer=tok)
    infer.input_video_stream(frames[:5])
    infer.input_query_stream([{"role": "user", "time": 1.0, "content": "what is the person doing now"}])
    loop = _oracle_loop(arch, w, infer, frames[:5], stream_end_prob_threshold=1.0)
    loop.input_query_stream([(1.0, lambda role: tok.apply_chat_template([{"role": "user", "content": "what is the person doing now"}],
                                                                        add_stream_query_prompt=role == "stream", add_stream_prompt=True))])
    # step both loops to just after the query turn
    for lp in (infer, loop):
        for _ in range(2):
            lp._encode_frame()
            lp.video_time += 0.5
        lp._encode_query()
    gap = loop.top2_gaps[-1][1]
    print("query turn: oracle top-2 gap", gap, "ids", loop.last_ids.tolist(), infer.last_ids.view(-1).tolist())
    assert infer.last_role == "user" and infer.past_key_values.length == len(loop.cache)
    if gap > 0.05:
        assert infer.last_ids.view(-1).tolist() == loop.last_ids.tolist()
    a, b = infer._encode_frame(), loop._encode_frame()
    assert abs(a["informative_score"] - b["informative_score"]) < TOL and abs(a["relevance_score"] - b["relevance_score"]) < TOL


def test_joint_embed_scatters_frames_at_placeholders(setup):
    """LiveMixin.joint_embed (models/modeling_live.py:35-48): ids are embedded (clamped to the vocabulary) and the positions
    holding config.v_placeholder_id are overwritten with visual_embed(frames), in order."""
    arch, w, model, tok, frames = setup
    v = model.config.v_placeholder_id
    assert v == tok.convert_tokens_to_ids("<image>") and v is not None
    ids = torch.tensor([[7, 8] + [v] * 49 + [9] + [v] * 49 + [10, arch.vocab + 3]], device="cuda")
    out = model.joint_embed(ids, frames[:2].cuda())
    assert out.shape == (1, ids.shape[1], arch.hidden)
    emb = model.visual_embed(frames[:2].cuda())
    table = model.get_input_embeddings().weight
    assert torch.equal(out[0, 2:51], emb[:49]) and torch.equal(out[0, 52:101], emb[49:])
    assert torch.equal(out[0, :2], table[torch.tensor([7, 8], device="cuda")]) and torch.equal(out[0, 51], table[9])
    assert torch.equal(out[0, -1], table[arch.vocab - 1])                      # clamp(max=vocab_size-1)
    wd = {k: t.cuda() for k, t in w.items()}
    ref = R.visual_embed(wd, arch, R.preprocess_frames(frames[:2]).bfloat16().float().cuda())
    assert (out[0, 2:51].float() - ref[:49]).abs().max() < TOL
    # forward() on ids + frames = the training-style call of the reference (video_head_live_llava_qwen.py:135-137)
    o = model(input_ids=ids[:, :51], frames=frames[:1].cuda(), use_cache=True, return_dict=True, logits_to_keep="none")
    assert o.informative_logits.shape == (1, 51, 2) and o.past_key_values.get_seq_length() == 51


def test_merged_lora_matches_unmerged_forward():
    """Row f2: the reference runs the base checkpoint with an UNMERGED peft adapter (models/modeling_live.py:117-123; r = 16,
    alpha = 32 on all seven projections of every decoder layer); here the adapter is folded into the bf16 matrices at load.
    The merged CUDA model must match the unmerged forward (fp32 oracle on W + (alpha/r) B A, nothing rounded after the merge)
    at the north-star tolerance, and the adapter must matter (scores move far more than the tolerance when it is dropped)."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.set_grad_enabled(False)
    from mmduet_b200 import build_model_and_tokenizer
    from mmduet_b200.config import ModelConfig
    arch = A.SMALL
    w = R.make_weights(arch, seed=61)
    g = torch.Generator().manual_seed(62)
    lora, w_eff = {}, dict(w)
    for i in range(arch.layers):
        for mod in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"):
            key = f"model.layers.{i}.{mod}"
            n, k = w[key + ".weight"].shape
            a = (torch.randn(16, k, generator=g) * k ** -0.5).bfloat16().float()
            b = (torch.randn(n, 16, generator=g) * 0.02).bfloat16().float()
            lora[f"base_model.model.{key}.lora_A.weight"], lora[f"base_model.model.{key}.lora_B.weight"] = a, b
            w_eff[key + ".weight"] = w[key + ".weight"] + 2.0 * (b @ a)
    cfg = ModelConfig.from_any(arch)
    merged, _ = build_model_and_tokenizer(state_dict=w, lora_state_dict=lora, lora_r=16, lora_alpha=32, model_config=cfg, device="cuda:0",
                                          max_context=1024)
    base, _ = build_model_and_tokenizer(state_dict=w, model_config=cfg, device="cuda:0", max_context=1024)
    frames = R.synthetic_frames(6, seed=63)
    px = R.preprocess_frames(frames).bfloat16().float().cuda()
    wd = {k: v.cuda() for k, v in w_eff.items()}
    from oracle import parity as P
    ref = P.oracle_stream(wd, arch, px, list(range(3, 20)), frames_per_pass=1)

    def run(model):
        emb = model.visual_embed(frames.cuda())
        st, L, sc = model.decoder.new_stream(), 0, []
        for f in range(6):
            o = model.decoder.step([dict(storage=st, past=L, ids=list(range(3, 20)) if f == 0 else [], frames=emb[f * 49:(f + 1) * 49])])
            L = o["views"][0].length
            sc.append(o["scores"][0])
        return torch.stack(sc)
    s_merged, s_base = run(merged), run(base)
    err = (s_merged - ref["scores"]).abs().max().item()
    moved = (s_base - ref["scores"]).abs().max().item()
    print("merged LoRA vs unmerged oracle", err, "| base model without the adapter", moved)
    assert err < TOL, err
    assert moved > 5 * TOL, "the synthetic adapter does not change the scores enough to test anything"


def test_loop_with_a_real_hf_tokenizer(tmp_path):
    """Row f2: a genuine PreTrainedTokenizerFast (built offline: byte-level BPE, ChatML specials) loaded through
    build_live_tokenizer_and_update_config drives LiveInferForBenchmark — system prompt, a user query, a generated response —
    and the CUDA loop agrees with the oracle loop fed the same token ids."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.set_grad_enabled(False)
    from mmduet_b200 import build_model_and_tokenizer
    from mmduet_b200.arguments_live import LiveTestArguments
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.inference import LiveInferForBenchmark, template_ids
    from tests.test_ingestion_cpu import offline_tokenizer_dir
    arch = A.TINY
    w = R.make_weights(arch, seed=71)
    d = offline_tokenizer_dir(str(tmp_path / "tok"))
    model, tok = build_model_and_tokenizer(state_dict=w, llm_pretrained=d, model_config=ModelConfig.from_any(arch), device="cuda:0", max_context=2048)
    assert type(tok).__name__ in ("PreTrainedTokenizerFast", "TokenizersBackend") or hasattr(tok, "backend_tokenizer")
    assert model.config.eos_token_id == tok.eos_token_id and model.config.v_placeholder_id == tok.convert_tokens_to_ids("<image>")
    frames = R.synthetic_frames(6, seed=72)
    args = LiveTestArguments(frame_fps=2, system_prompt="You watch.", stream_end_prob_threshold=1.0)
    infer = LiveInferForBenchmark(args, model=model, tokenizer=tok)
    infer.input_video_stream(frames)
    infer.input_query_stream([{"role": "user", "time": 1.0, "content": "hi?"}])
    resp = infer.inference()
    assert resp == [{"time": 1.0, "content": "hi?", "role": "user"}]
    loop = R.LiveLoopOracle(w, arch, start_ids=template_ids(tok, [{"role": "system", "content": "You watch."}]),
                            stream_prompt_ids=template_ids(tok, [{}], add_stream_prompt=True),
                            stream_generation_ids=template_ids(tok, [{}], add_stream_generation_prompt=True), eos_token_id=tok.eos_token_id,
                            frame_fps=2, stream_end_prob_threshold=1.0, max_new_tokens=4)
    loop.input_video_stream(R.preprocess_frames(frames).bfloat16().float())
    loop.input_query_stream([(1.0, lambda role: template_ids(tok, [{"role": "user", "content": "hi?"}], add_stream_query_prompt=role == "stream",
                                                             add_stream_prompt=True))])
    loop.inference()
    a = np.array([[x["informative_score"], x["relevance_score"]] for x in infer.debug_data_list])
    b = np.array([[x["informative_score"], x["relevance_score"]] for x in loop.debug_data_list])
    assert np.abs(a - b).max() < TOL and infer.past_key_values.length == len(loop.cache)
    # a response through the real tokenizer's decode()
    infer.inplace_output_ids = torch.zeros(1, 4, device=infer.device, dtype=torch.long)
    text = infer._generate_response()
    assert isinstance(text, str) and infer.last_role == "assistant"


def test_many_videos_leave_no_pages_or_memory_behind():
    """Production soak in miniature: 18 videos in a row (three passes over six videos of different length, half of them with a
    user query) through ONE LiveInferForBenchmark, responses generated and assistant turns rolled back, reset() between videos
    as test/inference.py:347-349 does: every KV page is back in the pool after each reset, and the third pass leaves exactly
    the allocator footprint and the results of the second (nothing accumulates, nothing depends on history)."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.set_grad_enabled(False)
    from mmduet_b200 import build_model_and_tokenizer
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.inference import LiveInferForBenchmark
    arch = A.SMALL
    w = R.make_weights(arch, seed=92)
    model, tok = build_model_and_tokenizer(state_dict=w, model_config=ModelConfig.from_any(arch), device="cuda:0", max_context=2048,
                                           kv_pages=96)
    dec = model.decoder
    n_free = len(dec._free)
    infer = LiveInferForBenchmark(_args(stream_end_score_sum_threshold=1.5, remove_assistant_turns=True), model=model, tokenizer=tok)
    infer.inplace_output_ids = torch.zeros(1, 6, device=infer.device, dtype=torch.long)
    videos = [R.synthetic_frames(10 + v % 3, seed=100 + v) for v in range(6)]
    mem, results, n_resp = [], [], 0
    for epoch in range(3):
        for v, frames in enumerate(videos):
            infer.reset()
            assert len(dec._free) == n_free, f"pass {epoch} video {v}: {n_free - len(dec._free)} pages still held after reset()"
            infer.set_fps(fps=2)
            infer.input_video_stream(frames)
            if v % 2:
                infer.input_query_stream([{"role": "user", "time": 1.0, "content": "what is happening"}])
            turns = infer.inference()
            infer.last_turns = turns
            n_resp += sum(t["role"] == "assistant" for t in turns)
            assert len(infer.debug_data_list) == len(frames)
            torch.cuda.synchronize()
            del turns
            torch.empty(1, device="cuda:0")                 # an allocation makes the caching allocator retire completed frees
            mem.append(torch.cuda.memory_allocated())
            results.append(([(t["role"], t["time"], t["content"]) for t in infer.last_turns], [d["informative_score"] for d in infer.debug_data_list]))
    infer.reset()
    assert len(dec._free) == n_free
    assert n_resp > 0                                     # the soak did generate and roll back
    assert mem[12:] == mem[6:12], mem
    assert results[12:] == results[6:12] == results[:6]
