"""GPU parity tests ON THE BENCHMARKED CONFIGURATION: full architecture (SigLIP-so400m + Qwen2-7B shapes, random init), the
decoder branches bench.py's default actually runs (40 frames = 1960/1992 tokens per pass: CTA-pair q/k/v, o, down with 3
split-K planes, interleaved swap-AB SwiGLU) and the >= 2048-token branch (CTA-pair pairwise SwiGLU), and the whole
BASELINE.json configs[1] stream (120 frames, 32-token prefix, 5.9k context) with the decision rule at the oracle's
80th-percentile threshold.  The checker is oracle/restate.py in fp32 on the GPU (oracle of record).

Tolerances (BASELINE.json north_star): max-abs 2e-2 on frame embeddings and on scores, identical threshold-crossing frames."""
import numpy as np
import pytest
import torch

from oracle import arch as A
from oracle import parity as P
from oracle import restate as R

pytestmark = pytest.mark.gpu
TOL = 2e-2
PREFIX = list(range(100, 132))


@pytest.fixture(scope="module")
def full():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_grad_enabled(False)
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.engine import DecoderEngine, VisionEngine
    dev = torch.device("cuda:0")
    arch = A.FULL
    w = R.make_weights(arch, seed=1234, device=dev, generate_on_device=True, include_lm_head=False)
    cfg = ModelConfig.from_any(arch)
    vis = VisionEngine(cfg, w, dev)
    dec = DecoderEngine(cfg, w, dev, max_context=8192, max_tokens=3072)
    yield arch, w, vis, dec, dev
    del vis, dec, w
    torch.cuda.empty_cache()


def _frame_end_rows(P0, k, n=49):
    return [P0 + n * (j + 1) - 1 for j in range(k)]


def _oracle_pass(w, arch, x, rows, cache=None):
    out = R.model_forward(w, arch, x, cache if cache is not None else R.KVCache(arch.layers))
    r = torch.tensor(rows, device=x.device)
    return torch.stack([out["informative_logits"][r].softmax(-1)[:, 1], out["relevance_logits"][r].softmax(-1)[:, 1]], 1)


@pytest.mark.parametrize("k,with_prefix", [(40, False), (40, True), (60, False), (21, True)])
def test_big_decoder_pass_matches_oracle(full, k, with_prefix):
    """One k-frame decoder pass at the full architecture: 40 frames = 1960 (1992 with the prefix) tokens is bench.py's
    default pass (M >= 1024: CTA-pair EPI_F32 q/k/v, o, down x 3 planes; M > 1024: EPI_T_SWIGLU_IL); 60 frames = 2940
    tokens takes the M >= 2048 branch (EPI_SWIGLU_PAIR); 21 frames + prefix = 1061 sits just above the switch.  Scores at
    every frame end vs the fp32 oracle, and vs the same frames fed one per step (the k = 1 path the reference runs)."""
    arch, w, vis, dec, dev = full
    g = torch.Generator(device=dev).manual_seed(100 + k)
    fe = (torch.randn(k * 49, arch.hidden, generator=g, device=dev) * 1.14).bfloat16()
    prefix = PREFIX if with_prefix else []
    P0 = len(prefix)
    rows = _frame_end_rows(P0, k)
    pre = R.embed_tokens(w, torch.tensor(prefix, device=dev, dtype=torch.long)) if prefix else torch.zeros(0, arch.hidden, device=dev)
    ref = _oracle_pass(w, arch, torch.cat([pre, fe.float()]), rows)
    st = dec.new_stream()
    out = dec.step([dict(storage=st, past=0, ids=prefix, frames=fe, score_rows=rows)], score="frame_ends")
    got = out["scores"].float()
    assert out["views"][0].length == P0 + 49 * k
    err = (got - ref).abs().max().item()
    # the k = 1 path over the same tokens
    st1 = dec.new_stream()
    L, one = 0, []
    for f in range(k):
        o = dec.step([dict(storage=st1, past=L, ids=prefix if f == 0 else [], frames=fe[f * 49:(f + 1) * 49])])
        L = o["views"][0].length
        one.append(o["scores"][0])
    one = torch.stack(one).float()
    err1 = (one - ref).abs().max().item()
    diff = (one - got).abs().max().item()
    print(f"k={k} prefix={P0}: pass-vs-oracle {err:.4f}  k1-vs-oracle {err1:.4f}  pass-vs-k1 {diff:.4f}")
    st.release()
    st1.release()
    assert err < TOL, f"{k}-frame pass scores vs oracle {err}"
    assert err1 < TOL, f"single-frame steps vs oracle {err1}"
    assert diff < TOL, f"{k}-frame pass vs single-frame steps {diff}"


def test_two_big_passes_with_history(full):
    """Second 40-frame pass on top of a 1992-token history (what passes 2 and 3 of the bench stream are): the paged
    tcgen05 attention front end with 1960 query tokens over ~4k keys, against the oracle with its cache."""
    arch, w, vis, dec, dev = full
    g = torch.Generator(device=dev).manual_seed(7)
    k = 40
    fe = (torch.randn(2 * k * 49, arch.hidden, generator=g, device=dev) * 1.14).bfloat16()
    cache = R.KVCache(arch.layers)
    pre = R.embed_tokens(w, torch.tensor(PREFIX, device=dev, dtype=torch.long))
    ref0 = _oracle_pass(w, arch, torch.cat([pre, fe[:k * 49].float()]), _frame_end_rows(32, k), cache)
    ref1 = _oracle_pass(w, arch, fe[k * 49:].float(), _frame_end_rows(0, k), cache)
    st = dec.new_stream()
    o0 = dec.step([dict(storage=st, past=0, ids=PREFIX, frames=fe[:k * 49], score_rows=_frame_end_rows(32, k))], score="frame_ends")
    o1 = dec.step([dict(storage=st, past=o0["views"][0].length, ids=[], frames=fe[k * 49:], score_rows=_frame_end_rows(0, k))],
                  score="frame_ends")
    e0, e1 = (o0["scores"] - ref0).abs().max().item(), (o1["scores"] - ref1).abs().max().item()
    print("pass 0", e0, "pass 1", e1)
    st.release()
    assert e0 < TOL and e1 < TOL, (e0, e1)


def test_two_streams_batched_pass_matches_solo(full):
    """configs[3]-style step: two videos in one decoder pass (2 x 24 frames = 2352 tokens, the >= 2048 branch); each
    stream's scores against the oracle run on that stream alone."""
    arch, w, vis, dec, dev = full
    g = torch.Generator(device=dev).manual_seed(9)
    k = 24
    fe = (torch.randn(2, k * 49, arch.hidden, generator=g, device=dev) * 1.14).bfloat16()
    rows = _frame_end_rows(0, k)
    refs = [_oracle_pass(w, arch, fe[i].float(), rows) for i in range(2)]
    sts = [dec.new_stream(), dec.new_stream()]
    out = dec.step([dict(storage=sts[i], past=0, ids=[], frames=fe[i], score_rows=rows) for i in range(2)], score="frame_ends")
    got = out["scores"].view(2, k, 2)
    errs = [(got[i] - refs[i]).abs().max().item() for i in range(2)]
    print("batched two-stream pass vs solo oracle", errs)
    for s in sts:
        s.release()
    assert max(errs) < TOL, errs


def test_configs1_stream_parity_and_crossings(full):
    """BASELINE.json configs[1] end to end: 120 synthetic frames through SigLIP + projector + pooling, 32-token prefix, three
    40-frame decoder passes (bench.py's default) AND 120 single-frame steps, against the fp32 oracle stream: frame
    embeddings, all 240 scores, and identical threshold-crossing frames at the oracle's 80th-percentile informative score."""
    arch, w, vis, dec, dev = full
    T = 120
    from mmduet_b200.random_init import synthetic_frames
    frames = synthetic_frames(T, seed=1, device=dev)                      # the frames bench.py streams on rank 0
    px = R.preprocess_frames(frames).bfloat16().float()
    ref = P.oracle_stream(w, arch, px, PREFIX, frames_per_pass=40)
    emb32 = vis.visual_embed(frames, normalize=True, out_dtype=torch.float32)
    emb = vis.visual_embed(frames, normalize=True)
    assert torch.equal(emb, emb32.bfloat16())

    def run(k):
        st, L, sc = dec.new_stream(), 0, []
        for f0 in range(0, T, k):
            nf = min(k, T - f0)
            p0 = 32 if f0 == 0 else 0
            o = dec.step([dict(storage=st, past=L, ids=PREFIX if f0 == 0 else [], frames=emb[f0 * 49:(f0 + nf) * 49],
                               score_rows=_frame_end_rows(p0, nf))], score="frame_ends")
            L = o["views"][0].length
            sc.append(o["scores"])
        assert L == 32 + 49 * T
        st.release()
        return torch.cat(sc, 0)

    s40, s1 = run(40), run(1)
    rep40 = P.parity_report(ref, s40, emb32)
    rep1 = P.parity_report(ref, s1)
    print("configs[1] k=40:", {k: v for k, v in rep40.items() if not k.startswith("crossings_")})
    print("configs[1] k=1 :", {k: v for k, v in rep1.items() if not k.startswith("crossings_")})
    print("k=40 vs k=1 scores", (s40 - s1).abs().max().item())
    assert rep40["emb_maxabs"] < TOL, rep40["emb_maxabs"]            # values before the final bf16 rounding (see test_gpu_parity)
    assert rep40["score_maxabs"] < TOL and rep1["score_maxabs"] < TOL, (rep40["score_maxabs"], rep1["score_maxabs"])
    assert rep40["n_crossings"] == 24
    for rep in (rep40, rep1, P.parity_report(ref, s40, head=1)):
        # identical crossing frames wherever the oracle score is further from the threshold than the measured error ...
        assert rep["flips_outside_noise"] == [], rep
        # ... and identical, full stop, at the threshold in the widest gap around the 80th percentile, whenever that gap
        # is wider than the error of the frames next to it
        wg = rep["widest_gap_near_quantile"]
        assert wg["crossings_match"] or wg["min_margin"] <= rep["score_maxabs"], rep
