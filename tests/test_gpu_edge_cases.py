"""GPU parity on the edge cases of the per-frame path: ragged multi-stream decoder passes (fresh stream + stream with history
+ single-token generation step in ONE pass), passes that end exactly on / one past a KV page boundary, one-frame and zero-frame
videos, the largest pass the engine was sized for, and loud refusals one past every limit.  Checker: oracle/restate.py (fp32)
run on each stream alone.  Tolerances as everywhere: 2e-2 on scores (BASELINE.json north_star); lm logits 6e-2 (bf16 lm_head)."""
import numpy as np
import pytest
import torch

from oracle import arch as A
from oracle import restate as R

pytestmark = pytest.mark.gpu
TOL = 2e-2
PREFIX = list(range(3, 35))


@pytest.fixture(scope="module")
def small():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_grad_enabled(False)
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.engine import DecoderEngine, VisionEngine
    dev = torch.device("cuda:0")
    arch = A.SMALL
    w = R.make_weights(arch, seed=21)
    wd = {k: v.to(dev) for k, v in w.items()}
    cfg = ModelConfig.from_any(arch)
    vis = VisionEngine(cfg, w, dev)
    dec = DecoderEngine(cfg, w, dev, max_context=1024, max_tokens=320, n_pages=48, max_lm_rows=3)
    return arch, wd, vis, dec, dev


def _frames(arch, n, seed, dev):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n * 49, arch.hidden, generator=g) * 1.1).bfloat16().to(dev)


def _oracle(wd, arch, cache, ids, fe):
    """one oracle pass over [ids | frame tokens] appended to `cache`; returns (scores at every row [n,2], lm logits of the last row)"""
    dev = fe.device if fe is not None else next(iter(wd.values())).device
    parts = []
    if ids:
        parts.append(R.embed_tokens(wd, torch.tensor(ids, device=dev, dtype=torch.long)))
    if fe is not None and fe.shape[0]:
        parts.append(fe.float())
    out = R.model_forward(wd, arch, torch.cat(parts), cache, want_lm_logits=True)
    sc = torch.stack([out["informative_logits"].softmax(-1)[:, 1], out["relevance_logits"].softmax(-1)[:, 1]], 1)
    return sc, out["logits"][-1]


def test_ragged_three_stream_pass(small):
    """One decoder pass over three streams of different shape: A fresh (32-token prefix + 3 frames, scores at the 3 frame ends),
    B with 130 tokens of history + 1 frame, C with 81 tokens of history + ONE text token (a generation step, lm logits read).
    Every stream against the oracle run on that stream alone."""
    arch, wd, vis, dec, dev = small
    fa, fb, fc = _frames(arch, 3, 1, dev), _frames(arch, 3, 2, dev), _frames(arch, 1, 3, dev)
    # histories, on both sides
    ca, cb, cc = R.KVCache(arch.layers), R.KVCache(arch.layers), R.KVCache(arch.layers)
    _oracle(wd, arch, cb, PREFIX, fb[:98])
    _oracle(wd, arch, cc, PREFIX, fc)
    sa, sb, sc_ = dec.new_stream(), dec.new_stream(), dec.new_stream()
    dec.step([dict(storage=sb, past=0, ids=PREFIX, frames=fb[:98])])
    dec.step([dict(storage=sc_, past=0, ids=PREFIX, frames=fc)])
    assert (sb.length, sc_.length) == (130, 81)
    ref_a, _ = _oracle(wd, arch, ca, PREFIX, fa)
    ref_b, _ = _oracle(wd, arch, cb, [], fb[98:])
    ref_c, lm_c = _oracle(wd, arch, cc, [77], None)
    out = dec.step([dict(storage=sa, past=0, ids=PREFIX, frames=fa, score_rows=[32 + 48, 32 + 97, 32 + 146]),
                    dict(storage=sb, past=130, ids=[], frames=fb[98:], score_rows=[48]),
                    dict(storage=sc_, past=81, ids=[77], score_rows=[0])], score="frame_ends", lm="last")
    got = out["scores"].float()
    assert got.shape == (5, 2)
    want = torch.cat([ref_a[[80, 129, 178]], ref_b[[48]], ref_c[[0]]])
    err = (got - want).abs().max().item()
    assert err < TOL, err
    assert [v.length for v in out["views"]] == [32 + 147, 179, 82]
    lm = out["lm_logits"].float()
    assert lm.shape[0] == 3
    assert (lm[2] - lm_c).abs().max().item() < 6e-2
    assert int(lm[2].argmax()) == int(lm_c.argmax()) or (lm_c.topk(2).values.diff().abs().item() < 6e-2)
    for s in (sa, sb, sc_):
        s.release()


@pytest.mark.parametrize("past", [63, 64, 65, 127, 128])
def test_pass_around_a_page_boundary(small, past):
    """History ending one before / on / one past a 64-token KV page boundary, then a frame that crosses the next one."""
    arch, wd, vis, dec, dev = small
    fe = _frames(arch, 1, 10 + past, dev)
    ids = [5 + (i % 200) for i in range(past)]
    cache = R.KVCache(arch.layers)
    _oracle(wd, arch, cache, ids, None)
    ref, _ = _oracle(wd, arch, cache, [], fe)
    st = dec.new_stream()
    dec.step([dict(storage=st, past=0, ids=ids)], score="none")
    out = dec.step([dict(storage=st, past=past, ids=[], frames=fe)])
    assert (out["scores"][0].float() - ref[48]).abs().max().item() < TOL
    assert st.length == past + 49 and len(st.pages) == (past + 49 + 63) // 64
    st.release()


def test_pass_sizes_around_the_workspace_and_the_context_limit(small):
    """A pass of exactly max_tokens rows, then one row more (the engine regrows its workspace: same numbers), both against the
    oracle; a pass that would exceed max_context is refused before anything changes."""
    arch, wd, vis, dec, dev = small
    from mmduet_b200 import _lib
    n = dec.max_tokens                                    # 320 = 6 frames + 26 ids
    fe = _frames(arch, 6, 5, dev)
    for extra in (0, 1):
        ids = PREFIX[:n - 6 * 49] + [9] * extra
        cache = R.KVCache(arch.layers)
        ref, _ = _oracle(wd, arch, cache, ids, fe)
        st = dec.new_stream()
        rows = [len(ids) + 49 * (j + 1) - 1 for j in range(6)]
        out = dec.step([dict(storage=st, past=0, ids=ids, frames=fe, score_rows=rows)], score="frame_ends")
        assert (out["scores"].float() - ref[rows]).abs().max().item() < TOL
        assert st.length == n + extra
        st.release()
    assert dec.max_tokens == n + 1
    # context limit: max_context = 1024
    big = dec.new_stream()
    L = 0
    for _ in range(4):
        o = dec.step([dict(storage=big, past=L, ids=[7] * 250)], score="none")
        L = o["views"][0].length
    free = len(dec._free)
    with pytest.raises(_lib.MmdError, match="max_context"):
        dec.step([dict(storage=big, past=L, ids=[7] * 25)], score="none")
    assert big.length == 1000 and len(dec._free) == free     # the refused pass left the stream and the pool where they were
    o = dec.step([dict(storage=big, past=L, ids=[7] * 24)], score="none")   # exactly max_context fits
    assert o["views"][0].length == 1024
    big.release()


def test_zero_and_one_frame_videos(small):
    """visual_embed of no frames is an empty [0, H] tensor (torch.cat semantics of the reference's batch loop,
    test/inference.py:203-206); a one-frame video gives one score pair equal to the oracle's; an empty decoder item is refused."""
    arch, wd, vis, dec, dev = small
    from mmduet_b200 import _lib
    empty = vis.visual_embed(torch.zeros(0, 3, 384, 384, dtype=torch.uint8, device=dev), normalize=True)
    assert tuple(empty.shape) == (0, arch.hidden)
    frames = R.synthetic_frames(1, seed=4)
    px = R.preprocess_frames(frames).bfloat16().float().to(dev)
    emb = vis.visual_embed(frames.to(dev), normalize=True)
    ref_emb = R.visual_embed(wd, arch, px)
    assert (emb.float() - ref_emb).abs().max().item() < TOL
    cache = R.KVCache(arch.layers)
    ref, _ = _oracle(wd, arch, cache, PREFIX, ref_emb)
    st = dec.new_stream()
    out = dec.step([dict(storage=st, past=0, ids=PREFIX, frames=emb)])
    assert (out["scores"][0].float() - ref[-1]).abs().max().item() < TOL
    with pytest.raises(_lib.MmdError, match="empty"):
        dec.step([dict(storage=st, past=st.length, ids=[], frames=empty)])
    assert st.length == 32 + 49
    st.release()


def test_loop_on_a_one_frame_and_an_empty_video():
    """The frame loop itself (LiveInferForBenchmark, the reference's driver surface) on degenerate videos: one frame -> one
    score record equal to the oracle loop's; no frames -> no records and no turns (test/inference.py:289-313 with an empty loop)."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.set_grad_enabled(False)
    from mmduet_b200 import build_model_and_tokenizer
    from mmduet_b200.arguments_live import LiveTestArguments
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.inference import LiveInferForBenchmark
    arch = A.SMALL
    w = R.make_weights(arch, seed=91)
    model, tok = build_model_and_tokenizer(state_dict=w, model_config=ModelConfig.from_any(arch), device="cuda:0", max_context=1024,
                                           kv_pages=64)
    args = LiveTestArguments(frame_fps=2, system_prompt="a b c d e f", stream_end_prob_threshold=1.0)
    frames = R.synthetic_frames(1, seed=5)
    infer = LiveInferForBenchmark(args, model=model, tokenizer=tok)
    infer.input_video_stream(frames)
    assert infer.inference() == []
    assert len(infer.debug_data_list) == 1
    loop = R.LiveLoopOracle(w, arch, start_ids=infer._start_ids.view(-1).tolist(),
                            stream_prompt_ids=infer._added_stream_prompt_ids.view(-1).tolist(),
                            stream_generation_ids=infer._added_stream_generation_ids.view(-1).tolist(),
                            eos_token_id=infer.eos_token_id, frame_fps=2, max_new_tokens=4, stream_end_prob_threshold=1.0)
    loop.input_video_stream(R.preprocess_frames(frames).bfloat16().float())
    assert loop.inference() == []
    for k in ("informative_score", "relevance_score"):
        assert abs(infer.debug_data_list[0][k] - loop.debug_data_list[0][k]) < TOL
    infer.reset()
    infer.input_video_stream(frames[:0])
    assert infer.inference() == [] and infer.debug_data_list == []
    assert infer.session.context_len == 0


@pytest.mark.parametrize("n_stages", [2, 3])
def test_layer_pipeline_stages_equal_the_whole_engine(small, n_stages):
    """The decoder split by layers into stages (parallel.LayerPipeline runs one per GPU; here chained on one GPU through the
    same step(resid_in=, resid_out=) interface): scores BIT-IDENTICAL to the whole engine over a multi-pass stream with a
    prefix, a big pass (appended precise rows), single-frame passes (every row hi+lo) and a rollback; lm logits too."""
    arch, wd, vis, dec, dev = small
    from mmduet_b200.parallel import layer_ranges
    fe = _frames(arch, 8, 33, dev)
    stages = [dec.stage(r) for r in layer_ranges(arch.layers, n_stages)]
    assert stages[0].first_stage and stages[-1].last_stage and sum(b - a for a, b in layer_ranges(arch.layers, n_stages)) == arch.layers
    passes = [(PREFIX, fe[:5 * 49], [32 + 49 * (j + 1) - 1 for j in range(5)]),      # 277 tokens: appended precise rows
              ([], fe[5 * 49:6 * 49], [48]), ([], fe[6 * 49:7 * 49], [48]),            # single-frame passes
              ([11, 12, 13], None, [2])]                                               # a text turn (lm logits)
    whole, L = dec.new_stream(), 0
    streams, Ls = [s.new_stream() for s in stages], 0
    for ids, fr, rows in passes:
        n_rows = len(ids) + (0 if fr is None else fr.shape[0])
        ref = dec.step([dict(storage=whole, past=L, ids=ids, frames=fr, score_rows=rows)], score="frame_ends", lm="last")
        L = ref["views"][0].length
        resid = None
        for k, (stg, sst) in enumerate(zip(stages, streams)):
            item = dict(storage=sst, past=Ls, ids=ids, frames=fr, score_rows=rows) if k == 0 else \
                dict(storage=sst, past=Ls, n_rows=n_rows, score_rows=rows)
            out = stg.step([item], score="frame_ends", lm="last", resid_in=resid, resid_out=k + 1 < n_stages)
            resid = out.get("resid")
        Ls = out["views"][0].length
        assert Ls == L
        assert torch.equal(out["scores"], ref["scores"]) and torch.equal(out["head_logits"], ref["head_logits"])
        assert torch.equal(out["lm_logits"], ref["lm_logits"])
    # rollback on every stage: re-append the last single-frame pass on top of the view before it
    ref = dec.step([dict(storage=whole, past=32 + 6 * 49, ids=[], frames=fe[7 * 49:], score_rows=[48])], score="frame_ends")
    resid = None
    for k, (stg, sst) in enumerate(zip(stages, streams)):
        item = dict(storage=sst, past=32 + 6 * 49, ids=[], frames=fe[7 * 49:], score_rows=[48]) if k == 0 else \
            dict(storage=sst, past=32 + 6 * 49, n_rows=49, score_rows=[48])
        out = stg.step([item], score="frame_ends", resid_in=resid, resid_out=k + 1 < n_stages)
        resid = out.get("resid")
    assert torch.equal(out["scores"], ref["scores"])
    from mmduet_b200 import _lib
    with pytest.raises(_lib.MmdError, match="resid_in"):
        stages[1].step([dict(storage=streams[1], past=0, n_rows=4)])          # a later stage needs the residual stream
    whole.release()
    for sst in streams:
        sst.release()
