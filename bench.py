#!/usr/bin/env python
"""bench.py — streamed frames/s of MMDuet's per-frame hot path (encode + KV-append + score) on B200.

One "step" = one pass over one synthetic video stream of BASELINE.json configs[1]:
    SigLIP-so400m + Qwen2-7B LiveLlava, 2 fps x 60 s = 120 frames of 384x384, random-init weights, a 32-token system
    prefix on frame 0, per-frame KV append + informative/relevance heads, decision rule applied to the scores.
`value`  : frames/s with the uint8 frames already resident in HBM, no host sync inside the stream.
`e2e`    : frames/s through the reference-facing API (LiveInferForBenchmark.input_video_stream + .inference) with HOST
           frames: the H2D copy of the frames and a D2H read of the two scores after every frame are inside the timing.
N > 1    : one process per GPU (torchrun), every rank streams its own videos (weak scaling, no data-path collective).
--impl reference : the CPU restatement of the reference path (oracle/) on the host cores, bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "streamed_frames_per_sec"
UNIT = "frames/s"
N_FRAMES = 120
PREFIX_LEN = 32


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                mx = float(f[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            busy = [s for s in sm if s > 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle restatement of the reference path on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_weights(device_sd=None, seed=1234):
    """bf16 weights on the host with the reference's key names (copied from the GPU arm's weights when available)."""
    import torch
    if device_sd is not None:
        return {k: v.to("cpu") for k, v in device_sd.items()}
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.random_init import random_state_dict
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sd = random_state_dict(ModelConfig(), seed=seed, device=dev, include_lm_head=False)
    return {k: v.to("cpu") for k, v in sd.items()}


def torch_gpu_port_sample(sd_dev, frames_u8_dev, n_frames, enc_batch=32):
    """The same restatement run on the GPU through stock PyTorch (cuBLAS GEMMs, SDPA, eager elementwise ops, growing
    torch.cat KV cache, one decoder call per frame exactly like the reference loop): what the reference's own code path
    delivers on this B200 — the practical bar next to the CPU number.  Baseline leg only, never the product path.
    Returns seconds for `n_frames` frames (device-synchronised wall clock)."""
    import torch
    from oracle import arch as A
    from oracle import restate as R
    arch = A.FULL
    dev = frames_u8_dev.device
    wd = {k: (v if v.dtype == torch.bfloat16 else v.to(torch.bfloat16)) for k, v in sd_dev.items()}
    # SDPA backends: the cuDNN one re-plans for every new KV length (1.4 ms of host time per call, measured), which a
    # growing cache hits on every step; the memory-efficient / math backends are what the reference effectively gets.
    from torch.nn.attention import SDPBackend, sdpa_kernel
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad(), sdpa_kernel([SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]):
        embs = []
        for b0 in range(0, n_frames, enc_batch):
            px = R.preprocess_frames(frames_u8_dev[b0:min(b0 + enc_batch, n_frames)]).to(torch.bfloat16)
            embs.append(R.visual_embed(wd, arch, px))
        emb = torch.cat(embs)
        cache = R.KVCache(arch.layers)
        prefix = torch.arange(100, 100 + PREFIX_LEN, device=dev)
        for f in range(n_frames):
            pre = R.embed_tokens(wd, prefix) if f == 0 else torch.zeros(0, arch.hidden, dtype=torch.bfloat16, device=dev)
            out = R.model_forward(wd, arch, torch.cat([pre, emb[f * 49:(f + 1) * 49]]), cache, attn_impl="sdpa")
            _ = out["informative_logits"][-1].softmax(-1)[1].item(), out["relevance_logits"][-1].softmax(-1)[1].item()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def cpu_reference_sample(w_cpu, frames_u8_cpu, n_frames):
    """One bounded sample of the workload on the CPU: encode `n_frames` frames (one batch) and run `n_frames` per-frame
    decoder steps (32-token prefix on the first) with the informative/relevance heads — the reference's
    visual_embed + forward + score path as restated in oracle/restate.py, bf16 like every reference script (--bf16 true).
    Returns seconds."""
    import torch
    from oracle import arch as A
    from oracle import restate as R
    arch = A.FULL
    wd = {k: (v.to(torch.bfloat16) if v.dtype != torch.bfloat16 else v) for k, v in w_cpu.items()}
    t0 = time.perf_counter()
    with torch.no_grad():
        px = R.preprocess_frames(frames_u8_cpu[:n_frames]).to(torch.bfloat16)
        emb = R.visual_embed(wd, arch, px)
        cache = R.KVCache(arch.layers)
        prefix = torch.arange(100, 100 + PREFIX_LEN)
        for f in range(n_frames):
            pre = R.embed_tokens(wd, prefix) if f == 0 else torch.zeros(0, arch.hidden, dtype=torch.bfloat16)
            out = R.model_forward(wd, arch, torch.cat([pre, emb[f * 49:(f + 1) * 49]]), cache)
            _ = out["informative_logits"][-1].softmax(-1)[1].item(), out["relevance_logits"][-1].softmax(-1)[1].item()
    return time.perf_counter() - t0


def run_reference_arm(args):
    """--impl reference: the reference path's CPU implementation as restated in oracle/restate.py (kind "port": the reference's
    own classes need /root/reference and LLaVA-NeXT, neither of which exists on the GPU box), bf16 like every reference
    script, on all host cores.  One step = a bounded sample of the configs[1] stream: the first `ref_frames` frames — one
    encoder batch, then that many per-frame decoder steps, the first carrying the 32-token prefix — so the whole
    --steps K --warmup W run ends within a few minutes; the sample shrinks (8 -> 4 -> 2 frames) if the first step
    shows that it would not."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    n_sample = args.ref_frames
    w = cpu_reference_weights()
    from mmduet_b200.random_init import synthetic_frames
    frames = synthetic_frames(n_sample, seed=1, device="cuda" if torch.cuda.is_available() else "cpu").cpu()
    cpu_reference_sample(w, frames, 1)                                  # thread pools, oneDNN primitives
    t_first = cpu_reference_sample(w, frames, n_sample)
    while n_sample > 2 and t_first * (args.steps + max(args.warmup - 1, 0)) > 200.0:
        t_first *= (n_sample // 2) / n_sample
        n_sample //= 2
    for _ in range(max(args.warmup - 1, 0)):
        cpu_reference_sample(w, frames, n_sample)
    times = [cpu_reference_sample(w, frames, n_sample) for _ in range(args.steps)]
    total = sum(times)
    v = n_sample * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, extra={"sample": f"{n_sample} frames encoded in one batch + {n_sample} decoder frame steps per step"}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{n_sample} frames (encode + {n_sample} per-frame decoder steps + heads), bf16, torch CPU"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _encoder_batch():
    from mmduet_b200.engine import VisionEngine
    return VisionEngine.MAX_BATCH


def workload_config(args, extra=None):
    c = {"workload": "BASELINE.json configs[1]: SigLIP-so400m/14@384 + Qwen2-7B LiveLlava frame step, 2 fps x 60 s = 120 synthetic "
                     "384x384 frames per stream, 32-token prefix, per-frame KV append + informative/relevance heads, random-init",
         "frames_per_step": N_FRAMES, "encoder_batch": _encoder_batch(), "decoder_frames_per_pass": args.chunk,
         "final_context_tokens": PREFIX_LEN + N_FRAMES * 49, "streams_per_gpu": 1,
         "l2": "weights (16 GB) and activations exceed the 126 MB L2 every step; no explicit flush"}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def algorithmic_work(tag, cfg, M_dec, T_enc):
    """(bytes, flops) one launch of `tag` must move / perform (DESIGN.md §4 states the same formulas): weights + activations
    read/written once, 2*M*N*K FLOP.  The binding roof is whichever of bytes/HBM-peak and flops/tensor-peak is larger: the
    decoder's weight-streaming GEMMs are HBM-bound at M=49 and tensor-bound from ~4 frames per pass (M >= 207)."""
    H, I, D, Dm = cfg.hidden, cfg.mlp, cfg.vit_dim, cfg.vit_mlp
    nqkv = (cfg.q_heads + 2 * cfg.kv_heads) * cfg.head_dim
    Mv = T_enc * cfg.patches
    L_avg = (PREFIX_LEN + N_FRAMES * 49) / 2.0
    table = {
        "gate_up_swiglu": (2 * I * H * 2 + M_dec * H * 2 + M_dec * I * 2, 2.0 * M_dec * 2 * I * H),
        "down_proj": (H * I * 2 + M_dec * I * 2 + M_dec * H * 4, 2.0 * M_dec * H * I),
        "qkv_proj": (nqkv * H * 2 + M_dec * H * 2 + M_dec * nqkv * 4, 2.0 * M_dec * nqkv * H),
        "o_proj": (H * H * 2 + M_dec * H * 2 + M_dec * H * 4, 2.0 * M_dec * H * H),
        "fc1": (Mv * D * 2 + Dm * D * 2 + Mv * Dm * 2, 2.0 * Mv * Dm * D),
        "fc2": (Mv * Dm * 2 + Dm * D * 2 + 2 * Mv * D * 4, 2.0 * Mv * Dm * D),
        "qkv": (Mv * D * 2 + 3 * D * D * 2 + Mv * 3 * D * 2, 2.0 * Mv * 3 * D * D),
        "out_proj": (Mv * 2 * D * 2 + 2 * D * D * 2 + 2 * Mv * D * 4, 2.0 * Mv * D * 2 * D),
        "vit_attention": (Mv * 3 * D * 2 + Mv * 2 * D * 2, 4.0 * T_enc * cfg.vit_heads * cfg.patches * cfg.patches * cfg.vit_head_dim),
        # KV-append attention: average over the stream's passes (keys visible ~ half the final context)
        "kv_attention": (2 * L_avg * cfg.kv_heads * cfg.head_dim * 2 + 2 * M_dec * cfg.q_heads * cfg.head_dim * 2,
                         4.0 * M_dec * L_avg * cfg.q_heads * cfg.head_dim),
    }
    return table.get(tag)


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner must not land on stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    from mmduet_b200 import _lib
    from mmduet_b200.arguments_live import LiveTestArguments
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.inference import LiveInferForBenchmark
    from mmduet_b200.modeling_live import VideoHeadLiveLlavaQwenForCausalLM
    from mmduet_b200.random_init import random_state_dict, synthetic_frames
    from mmduet_b200.tokenization_live import SyntheticTokenizer

    cfg = ModelConfig()
    torch.manual_seed(1234)
    sd = random_state_dict(cfg, seed=1234, device=dev, include_lm_head=False)
    ctx_len = max(PREFIX_LEN + N_FRAMES * 49 + 64, (CONFIGS2_FRAMES * 49 + 64) if args.configs2 else 0)
    model = VideoHeadLiveLlavaQwenForCausalLM(cfg, sd, device=dev, max_context=ctx_len, max_step_tokens=PREFIX_LEN + 49 * max(args.chunk, 1))
    vis, dec = model.vision, model.decoder
    frames_dev = synthetic_frames(N_FRAMES, seed=1 + rank, device=dev)
    frames_host = frames_dev.cpu().pin_memory()
    prefix = list(range(100, 100 + PREFIX_LEN))
    if not args.keep_weights_for_cpu:
        pass
    threshold = 0.8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stream_pass():
        """value path: device-resident frames, no host sync inside; returns the [N_FRAMES,2] device score tensor."""
        emb = vis.visual_embed(frames_dev, normalize=True)
        st = dec.new_stream()
        L, scores = 0, []
        k = max(args.chunk, 1)
        for f0 in range(0, N_FRAMES, k):
            nf = min(k, N_FRAMES - f0)
            item = dict(storage=st, past=L, ids=prefix if f0 == 0 else [], frames=emb[f0 * 49:(f0 + nf) * 49])
            p = PREFIX_LEN if f0 == 0 else 0
            item["score_rows"] = [p + 49 * (j + 1) - 1 for j in range(nf)]
            out = dec.step([item], score="frame_ends")
            L = out["views"][0].length
            scores.append(out["scores"])
        sc = torch.cat(scores, 0)
        st.release()
        return sc

    def decide(sc_host):
        # LiveInferForBenchmark.inference's rule with score_heads='informative_score', single-frame threshold (strict >)
        return [i for i, s in enumerate(sc_host[:, 0].tolist()) if s > threshold]

    # ---- warm-up + a profiling pass that finds the dominant kernel ----
    for _ in range(max(args.warmup, 3)):
        stream_pass()
    barrier()
    _lib.profile_start("all", local)
    stream_pass()
    torch.cuda.synchronize()
    stage = _lib.profile_stop(local)
    stage_ms = {k: round(v[0], 3) for k, v in sorted(stage.items(), key=lambda kv: -kv[1][0])}
    known = {k: v for k, v in stage.items() if algorithmic_work(k, cfg, 49, _encoder_batch()) is not None}
    roof_tags = {}
    dominant = max(known.items(), key=lambda kv: kv[1][0])[0]

    # ---- timed region: value ----
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    _lib.profile_start(dominant, local)
    launches0 = _lib.launch_count(local)
    e0.record()
    last = None
    for _ in range(args.steps):
        last = stream_pass()
    e1.record()
    barrier()
    launches = _lib.launch_count(local) - launches0
    clocks = sampler.stop()
    dom = _lib.profile_stop(local)[dominant]
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    value = world * N_FRAMES * args.steps / (ms_max / 1e3)
    crossings = decide(last.cpu())

    # ---- timed region: e2e through the reference-facing API with host frames ----
    targs = LiveTestArguments(stream_end_prob_threshold=1.0, frame_fps=2, score_heads="informative_score")  # grounding-style: never generates
    tok = SyntheticTokenizer(cfg.vocab)
    infer = LiveInferForBenchmark(targs, model=model, tokenizer=tok)
    infer._start_ids = torch.tensor([prefix], device=dev)   # same 32-token prefix as the value path
    infer.frames_per_step = max(args.chunk, 1)
    lat = []

    def e2e_pass(record=False):
        infer.reset()
        infer.input_video_stream(frames_host)
        infer.session.step_ms = lat if record else None     # host wall clock of every frame pass (launch + score read-back)
        infer.inference()
        infer.session.step_ms = None
        return infer.debug_data_list

    for _ in range(2):
        e2e_pass()
    barrier()
    n_e2e = max(1, min(args.steps, 5))
    e0.record()
    for _ in range(n_e2e):
        dbg = e2e_pass()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * N_FRAMES * n_e2e / (t.item() / 1e3)
    infer.frames_per_step = 1      # latency mode: one frame per decoder pass, score read back after every frame
    e2e_pass()
    e2e_pass(record=True)
    lat_sorted = sorted(lat)
    # per-frame encode latency in live (one frame at a time) mode
    one = frames_dev[:1]
    for _ in range(3):
        vis.visual_embed(one, normalize=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        vis.visual_embed(one, normalize=True)
    e1.record()
    torch.cuda.synchronize()
    enc1_ms = e0.elapsed_time(e1) / 10
    # consistency of the two paths (same frames, same prefix): scores must agree
    e2e_scores = torch.tensor([[d["informative_score"], d["relevance_score"]] for d in dbg])
    path_diff = (e2e_scores - last.cpu()).abs().max().item() if rank == 0 else 0.0
    k1_scores = torch.tensor([[d["informative_score"], d["relevance_score"]] for d in infer.debug_data_list])   # the latency pass (k = 1)
    # one LIVE frame end to end: uint8 frame in pinned host memory -> H2D -> SigLIP + projector + pool -> decoder step -> scores on the host
    infer.reset()
    infer.frames_per_step = 1
    live = []
    for f in range(N_FRAMES):
        t0 = time.perf_counter()
        infer.input_video_stream(frames_host[f:f + 1])
        sc_live = infer._encode_frame()
        live.append((time.perf_counter() - t0) * 1e3)
    live_sorted = sorted(live[4:])
    # BASELINE configs[2] under the same clock: frame-parallel encoder -> exchange -> owner decodes (all ranks take part)
    infer.reset()
    c2 = None
    if args.configs2:
        need = CONFIGS2_FRAMES * 49 + 64
        if dec.max_context >= need:
            c2 = configs2_frame_parallel(vis, dec, cfg, dev, world, rank, args.chunk)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    work = algorithmic_work(dominant, cfg, 49 * max(args.chunk, 1), _encoder_batch())
    dom_ms, dom_n = dom
    roofline = None
    if work is not None and dom_n > 0:
        per_launch_s = dom_ms / dom_n / 1e3
        nbytes, nflops = work
        t_hbm, t_tc = nbytes / (pk["hbm_gbs"] * 1e9), nflops / (pk["bf16_tflops_sustained"] * 1e12)
        common = {"kernel": dominant, "traffic": ncu_traffic(dominant, max(args.chunk, 1)), "launches_timed": dom_n, "avg_launch_us": per_launch_s * 1e6,
                  "algorithmic_bytes_per_launch": nbytes, "algorithmic_flops_per_launch": nflops,
                  "share_of_stream": round(stage[dominant][0] / sum(v[0] for v in stage.values()), 3)}
        if t_hbm >= t_tc:
            ach = nbytes / per_launch_s / 1e9
            roofline = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                        "peak_source": pk["source"], **common}
        else:
            ach = nflops / per_launch_s / 1e12
            roofline = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / pk["bf16_tflops_sustained"], "peak_source": pk["source"] + " (sustained, kernel timed inside a long step)",
                        **common}
    # informative: the same roofline arithmetic for every kernel with a formula, from the profiling pass (not the timed region)
    roof_all = []
    tot_ms = sum(v[0] for v in stage.values())
    for tag, (ms_t, n_t) in sorted(stage.items(), key=lambda kv: -kv[1][0]):
        wk = algorithmic_work(tag, cfg, 49 * max(args.chunk, 1), _encoder_batch())
        if wk is None or n_t == 0:
            continue
        sec = ms_t / n_t / 1e3
        t_hbm, t_tc = wk[0] / (pk["hbm_gbs"] * 1e9), wk[1] / (pk["bf16_tflops_sustained"] * 1e12)
        if t_hbm >= t_tc:
            roof_all.append({"kernel": tag, "share": round(ms_t / tot_ms, 3), "bound": "hbm", "achieved_GBps": round(wk[0] / sec / 1e9, 1),
                             "frac": round(wk[0] / sec / 1e9 / pk["hbm_gbs"], 3)})
        else:
            roof_all.append({"kernel": tag, "share": round(ms_t / tot_ms, 3), "bound": "tensor", "achieved_TFLOPs": round(wk[1] / sec / 1e12, 1),
                             "frac": round(wk[1] / sec / 1e12 / pk["bf16_tflops_sustained"], 3)})
            if tag in ("vit_attention", "kv_attention"):
                # the attention kernels' real ceiling is MUFU.EX2 (16 lanes/clk/SM), not the tensor pipe: one exponential
                # per score = QK^T+PV flops / (4 * head_dim); peak at the maximum SM clock (conservative)
                dh = cfg.vit_head_dim if tag == "vit_attention" else cfg.head_dim
                exps = wk[1] / (4.0 * dh)
                mufu_peak = 16.0 * torch.cuda.get_device_properties(dev).multi_processor_count * (clocks.get("sm_max_mhz") or 1965.0) * 1e6
                roof_all[-1].update(mufu_bound_frac=round(exps / sec / mufu_peak, 3), exps_per_launch=int(exps))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(frames_host.numel()), "d2h_bytes_per_step": N_FRAMES * 8,
                    "steps": n_e2e},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "latency_ms": {"decoder_frames_per_pass": 1, "p50_frame_step": lat_sorted[len(lat_sorted) // 2], "p99_frame_step": lat_sorted[int(len(lat_sorted) * 0.99) - 1],
                           "single_frame_encode": enc1_ms,
                           "frame_step_roofline": live_step_roofline(cfg, lat_sorted[len(lat_sorted) // 2], pk),
                           "p50_live_frame": live_sorted[len(live_sorted) // 2], "p99_live_frame": live_sorted[int(len(live_sorted) * 0.99) - 1],
                           "note": "live frame = ONE uint8 frame in pinned host memory -> H2D -> SigLIP+projector+pool -> decoder KV-append + heads -> "
                           "two scores on the host, host wall clock per frame over a 120-frame stream (context grows to 5.9k); frame step = the "
                           "decoder part alone; encode = the encoder part alone"},
            "stage_ms_per_stream": stage_ms, "roofline_by_kernel": roof_all, "threshold_crossings": crossings[:16], "value_vs_e2e_score_maxdiff": path_diff}
    if c2 is not None:
        line["configs2_frame_parallel"] = c2
    if args.parity:
        try:
            emb32 = vis.visual_embed(frames_dev, normalize=True, out_dtype=torch.float32)
            line["parity"] = parity_block(sd, cfg, frames_dev, prefix, emb32, {max(args.chunk, 1): last, 1: k1_scores.to(dev)})
            del emb32
        except Exception as e:  # noqa: BLE001
            line["parity"] = {"failed": repr(e)}
    if args.cpu_baseline and world == 1:
        try:
            fr_dev = frames_host.to(dev)
            torch_gpu_port_sample(sd, fr_dev, 8)   # warm-up (cuBLAS heuristics, SDPA kernels)
            sec = torch_gpu_port_sample(sd, fr_dev, N_FRAMES)
            line["torch_gpu_baseline"] = {"value": N_FRAMES / sec, "unit": UNIT, "kind": "port",
                                          "sample": f"the oracle restatement on cuda:0 (PyTorch eager bf16, cuBLAS + SDPA (mem-efficient backend), encoder batch 32, "
                                                    f"{N_FRAMES} per-frame decoder steps as the reference loop does), inputs resident, {sec:.2f} s"}
            del fr_dev
        except Exception as e:  # noqa: BLE001
            line["torch_gpu_baseline"] = {"value": None, "unit": UNIT, "kind": "port", "sample": f"failed: {e!r}"}
        try:
            torch.set_num_threads(os.cpu_count() or 1)
            w_cpu = cpu_reference_weights(sd)
            n_s = args.cpu_frames
            fr = frames_host[:n_s].clone()
            cpu_reference_sample(w_cpu, fr, 1)  # warm-up (thread pools, oneDNN primitives)
            sec = cpu_reference_sample(w_cpu, fr, n_s)
            line["cpu_baseline"] = {"value": n_s / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{n_s} frames: one encode batch + {n_s} per-frame decoder steps + heads, bf16, torch CPU, {sec:.1f} s"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def live_step_roofline(cfg, p50_ms, pk):
    """The single-frame decoder step streams every decoder-layer weight once (49 tokens: HBM-bound): algorithmic bytes =
    layers x (q/k/v + o + gate/up + down) bf16 weights + the 49-row KV append, against the measured HBM peak."""
    H, I, dh = cfg.hidden, cfg.mlp, cfg.head_dim
    per_layer = ((cfg.q_heads + 2 * cfg.kv_heads) * dh * H + H * cfg.q_heads * dh + 2 * I * H + H * I) * 2
    nbytes = cfg.layers * per_layer
    ach = nbytes / (p50_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "algorithmic_bytes_per_step": int(nbytes), "floor_ms": nbytes / (pk["hbm_gbs"] * 1e9) * 1e3,
            "note": "whole step (about 250 kernel launches) against the weight stream alone; host wall clock p50"}


def parity_block(sd, cfg, frames_dev, prefix, emb32, scores_by_k):
    """The fp32 oracle (oracle/restate.py on bf16-rounded weights and pixels, run on the GPU, OUTSIDE every timed region) over
    the same 120-frame stream, as the checker of the scores the timed passes produced: north-star tolerances are max-abs
    2e-2 on frame embeddings (measured on the values before the final bf16 rounding; `emb_bf16_out_maxabs` is the rounded
    output, whose floor `emb_bf16_floor` is the rounding of the exact oracle values) and on scores, and identical
    threshold-crossing frames at the oracle's 80th-percentile informative score."""
    import torch
    from oracle import arch as A
    from oracle import parity as P
    from oracle import restate as R
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        w32 = {k: v.float() for k, v in sd.items()}
        px = R.preprocess_frames(frames_dev).bfloat16().float()
        t0 = time.perf_counter()
        ref = P.oracle_stream(w32, A.FULL, px, prefix, frames_per_pass=40)
        torch.cuda.synchronize()
        out = {"oracle": "oracle/restate.py, fp32 on bf16-rounded weights/pixels, on the GPU (stock PyTorch), 40 frames per causal pass",
               "oracle_seconds": round(time.perf_counter() - t0, 2), "tolerance": 2e-2}
        for k, sc in scores_by_k.items():
            rep = P.parity_report(ref, sc, emb32 if k == max(scores_by_k) else None)
            out[f"frames_per_pass_{k}"] = {kk: (round(v, 6) if isinstance(v, float) else v) for kk, v in rep.items()}
        main = out[f"frames_per_pass_{max(scores_by_k)}"]
        reps = [out[f"frames_per_pass_{k}"] for k in scores_by_k]
        out.update(emb_maxabs=main.get("emb_maxabs"), score_maxabs=max(r["score_maxabs"] for r in reps),
                   crossings_match=all(r["crossings_match"] for r in reps), min_margin=main["min_margin"],
                   flips_outside_noise=sum(len(r["flips_outside_noise"]) for r in reps),
                   crossings_match_widest_gap=all(r["widest_gap_near_quantile"]["crossings_match"] for r in reps),
                   min_margin_widest_gap=main["widest_gap_near_quantile"]["min_margin"])
        out["emb_bf16_out_maxabs"] = round(float((emb32.bfloat16().float() - ref["emb"]).abs().max()), 6)
        del w32, ref
        torch.cuda.empty_cache()
        return out
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


CONFIGS2_FRAMES = 600


def configs2_frame_parallel(vis, dec, cfg, dev, world, rank, chunk):
    """BASELINE.json configs[2]: one 5-minute video at 2 fps (600 frames); the ENCODER is sharded over the N ranks (contiguous
    frame ranges), every rank's projector/pool kernel stores its frame tokens straight into the owner's HBM (symmetric memory
    over NVLink, `PeerStoreEncoder`; NCCL send/recv if symmetric memory is unavailable), and rank 0 — the owner of the video's
    decoder stream — decodes all 600 frames in `chunk`-frame passes as the batches land, applying the YouCook2 decision rule
    (running score sum > 2, youcook2.sh:14) to the scores in frame order.  Responses are not generated: with
    remove_assistant_turns they do not change the context, so every later score is unaffected (stated in `config`).
    Timed on the device from before the first encode launch to the last score, max over ranks."""
    import torch
    import torch.distributed as dist
    from mmduet_b200.parallel import (FrameParallelEncoder, LayerPipeline, PeerStoreEncoder, encoder_batches, encoder_frame_range,
                                      layer_ranges)
    from mmduet_b200.random_init import synthetic_frames
    n, tpf, H = CONFIGS2_FRAMES, vis.tokens_per_frame, cfg.hidden
    k = max(chunk, 1)
    frames = synthetic_frames(n, seed=7, device=dev)              # the same video on every rank; each encodes its slice
    # The video's decoder stream is the serial term.  Ranks 0..S-1 hold it as a LAYER PIPELINE (stage s runs its layers of pass p
    # while stage s-1 runs pass p+1; parallel.LayerPipeline), ranks S..N-1 encode; the decoder ranks do not encode when there
    # are other ranks.  One GPU's decode of this video costs ~1.2x its encode, hence S ~ 0.55 N.
    S = 1 if world < 2 else min(world - 1, max(1, int(round(0.55 * world))))
    # Opt-in (MMD_CONFIGS2_ALL_BOTH=1, NOT the default): every rank is a decoder stage AND encodes its round-robin batches on a
    # side stream.  Measured 561 ms at N = 2 (default split: 602) and 303 ms at N = 4 (341), bit-identical scores, but the same
    # run HUNG at N = 8 (side-stream signal waits + NCCL point-to-point on 8 ranks; not diagnosed), so the default keeps
    # dedicated decoder and encoder ranks.
    both = world >= 2 and os.environ.get("MMD_CONFIGS2_ALL_BOTH") == "1"
    if both:
        S = world
    S = int(os.environ.get("MMD_CONFIGS2_STAGES", S))
    encoders = (list(range(world)) if both else list(range(S, world))) if world > 1 else [0]
    enc_stream = torch.cuda.Stream(device=dev) if both else None
    lo, hi = encoder_frame_range(n, encoders, rank)
    pipe = LayerPipeline(list(range(S)), H, dev) if S > 1 else None
    stage_ranges = layer_ranges(cfg.layers, S)
    stage_eng = dec.stage(stage_ranges[rank], max_tokens=49 * k + 64) if S > 1 and rank < S else None

    def encode_video():
        if enc_stream is None:
            return enc.encode(n, local_frames)
        enc_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(enc_stream):          # the encode launches run beside this rank's decoder stage
            return enc.encode(n, local_frames)
    passes = [dict(rows=min(k, n - f0) * tpf, f0=f0, nf=min(k, n - f0)) for f0 in range(0, n, k)]
    exchange = "none (1 GPU)"
    enc = None
    if world > 1:
        try:
            # batches of 40 frames dealt round-robin to the encoder ranks: the video becomes available front to back at the
            # encoders' aggregate rate, which is the order the decoder consumes it in
            enc = PeerStoreEncoder(lambda fr, dst: vis.visual_embed(fr, normalize=True, out=dst), tpf, H, max_frames=n, device=dev,
                                   owner=0, batch=40, encoders=encoders, assignment="round_robin")
            mine = encoder_batches(n, encoders, rank, 40, "round_robin")
            local_frames = torch.cat([frames[b0:b1] for b0, b1 in mine]) if mine else frames[:0]
            exchange = ("peer stores into the owner's HBM (symmetric memory over NVLink), 40-frame batches dealt round-robin to the "
                        "encoder ranks, per-batch signals, no collective")
        except Exception as e:  # noqa: BLE001
            enc = FrameParallelEncoder(lambda fr: vis.visual_embed(fr, normalize=True), tpf, H, device=dev, owner=0, batch=40, encoders=encoders)
            local_frames = frames[lo:hi]
            exchange = f"NCCL isend/irecv per 40-frame batch (symmetric memory unavailable: {type(e).__name__})"
    dec._ensure_ws(49 * k, 1)
    if dec.max_context < n * tpf:
        raise RuntimeError("decoder context too small for configs[2]")

    def owner_decode(tokens, ready):
        st, L, sc = dec.new_stream(), 0, []
        for f0 in range(0, n, k):
            nf = min(k, n - f0)
            if ready is not None:
                ready[f0 + nf - 1]()                                # stream-ordered wait until this pass's last frame has landed
            o = dec.step([dict(storage=st, past=L, ids=[], frames=tokens[f0 * tpf:(f0 + nf) * tpf],
                               score_rows=[tpf * (j + 1) - 1 for j in range(nf)])], score="frame_ends")
            L = o["views"][0].length
            sc.append(o["scores"])
        st.release()
        return torch.cat(sc, 0), L

    def pipelined_decode(tokens, ready):
        """this rank's stage of the layer pipeline over all passes; the scores end up on rank 0 (the video's owner)"""
        st, state = stage_eng.new_stream(), {"L": 0}

        def stage_fn(p, d, rin):
            f0, nf = d["f0"], d["nf"]
            rows = [tpf * (j + 1) - 1 for j in range(nf)]
            if pipe.is_first:
                if ready is not None:
                    ready[f0 + nf - 1]()
                item = dict(storage=st, past=state["L"], ids=[], frames=tokens[f0 * tpf:(f0 + nf) * tpf], score_rows=rows)
            else:
                item = dict(storage=st, past=state["L"], n_rows=nf * tpf, score_rows=rows)
            o = stage_eng.step([item], score="frame_ends", resid_in=rin, resid_out=not pipe.is_last)
            state["L"] = o["views"][0].length
            return o["scores"] if pipe.is_last else o["resid"]
        res = pipe.run(passes, stage_fn)
        st.release()
        sc = None
        if pipe.is_last:
            sc = torch.cat(res, 0)
            dist.send(sc, dst=0)
        if rank == 0:
            sc = torch.empty(n, 2, dtype=torch.float32, device=dev)
            dist.recv(sc, src=S - 1)
        return sc, state["L"]

    def one_pass():
        if world == 1:
            tokens = vis.visual_embed(frames, normalize=True)
            return owner_decode(tokens, None) + (tokens,)
        tokens, ready = encode_video()
        if rank >= S:
            return None, None, None
        sc, L = pipelined_decode(tokens, ready) if S > 1 else owner_decode(tokens, ready)
        if rank == 0:
            FrameParallelEncoder.wait_all(ready)
        if enc_stream is not None:
            torch.cuda.current_stream(dev).wait_stream(enc_stream)
        return sc, L, tokens

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    one_pass()
    barrier()
    e0, e1, e_enc = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record()
    for _ in range(reps):
        sc, L, tokens = one_pass()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    # the two halves separately (not overlapped), for the Amdahl statement: encode+exchange alone, owner decode alone
    barrier()
    e0.record()
    if world == 1:
        tokens = vis.visual_embed(frames, normalize=True)
    else:
        tokens, ready = encode_video()
        FrameParallelEncoder.wait_all(ready)
    e_enc.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e_enc)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    enc_ms = t.item()
    res = None
    if rank == 0:
        e0.record()
        sc2, _ = owner_decode(tokens, None)
        e1.record()
        torch.cuda.synchronize()
        dec_ms = e0.elapsed_time(e1)
        single = vis.visual_embed(frames, normalize=True)           # single-rank encode of the whole video: the bit-identity check
        ident = bool(torch.equal(single, tokens[:n * tpf]))
        s_host = sc[:, 0].double().cpu().tolist()
        acc, resp = 0.0, []
        for i, v in enumerate(s_host):                              # test/inference.py:296-298 with stream_end_score_sum_threshold = 2
            acc += v
            if acc > 2.0:
                resp.append(i)
                acc = 0.0
        res = {"workload": "BASELINE.json configs[2]: one 600-frame video (5 min @ 2 fps), encoder sharded over the ranks, frame tokens "
                           "exchanged to rank 0, which decodes the whole stream (final context 29.4k tokens) and applies the running-sum rule "
                           "(threshold 2, informative head); responses not generated (remove_assistant_turns: context-neutral)",
               "frames": n, "n_ranks": world, "n_encoder_ranks": len(encoders), "decoder_rank_also_encodes": world == 1 or both,
               "exchange": exchange,
               "n_decoder_stages": S, "decoder_layers_per_stage": [b_ - a_ for a_, b_ in stage_ranges],
               "decoder_exchange": "none (one decoder rank)" if S == 1 else
                                   f"layer pipeline: fp32 residual stream [{k * tpf}, {H}] handed from stage to stage per pass (NCCL isend/irecv, "
                                   f"{k * tpf * H * 4 / 1e6:.0f} MB), KV pages of a layer live on its stage only",
               "decoder_frames_per_pass": k,
               "ms": ms, "frames_per_s": n / (ms / 1e3), "encode_exchange_only_ms": enc_ms, "owner_decode_only_ms": dec_ms,
               "final_context_tokens": int(L), "responses": len(resp), "first_response_frames": resp[:8],
               "tokens_bit_identical_to_single_rank_encode": ident,
               "scores_equal_overlapped_vs_separate": bool(torch.equal(sc, sc2)),
               "limiter": "the owner's decoder stream is sequential (KV dependency): Amdahl's serial term is owner_decode_only_ms" if S == 1 else
                          f"one GPU's decode of the video (owner_decode_only_ms) is divided over {S} layer stages (+ {S - 1} passes of pipeline "
                          f"fill); {len(encoders)} encoder ranks run beside it"}
    del frames
    return res


def run_encoder16(args):
    """--workload encoder16 = BASELINE.json configs[0]: SigLIP-so400m/14@384 tower + mm_projector + pooling over 16 synthetic
    384x384 frames (no decoder); the reference's CPU path (oracle restatement, fp32 and bf16, SDPA-free eager attention as
    in oracle/restate.py) timed beside it on the host cores and used as the checker."""
    import torch
    from mmduet_b200 import _lib
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.engine import VisionEngine
    from mmduet_b200.random_init import random_state_dict, synthetic_frames
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cfg = ModelConfig()
    sd = {k: v for k, v in random_state_dict(cfg, seed=1234, device=dev, include_lm_head=False).items() if not k.startswith("model.layers")}
    vis = VisionEngine(cfg, sd, dev)
    T = 16
    frames = synthetic_frames(T, seed=1, device=dev)
    host = frames.cpu().pin_memory()
    for _ in range(max(args.warmup, 3)):
        vis.visual_embed(frames, normalize=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count(0)
    e0.record()
    for _ in range(args.steps):
        emb = vis.visual_embed(frames, normalize=True)
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count(0) - l0
    ms = e0.elapsed_time(e1) / args.steps
    e0.record()
    for _ in range(args.steps):
        emb_h = vis.visual_embed(host.to(dev, non_blocking=True), normalize=True)
        chk = emb_h[-1, :8].float().cpu()                           # D2H read of the step's result
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    line = {"metric": METRIC, "value": T / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[0]: SigLIP-so400m/14@384 vision tower + mm_projector + bilinear 27->7 pooling over "
                                   "16 synthetic 384x384 frames, random-init (encoder only, no decoder)", "frames_per_step": T,
                       "l2": "0.8 GB of weights + activations exceed L2 every step; no explicit flush"},
            "e2e": {"value": T / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(host.numel()), "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches), "clocks": clocks}
    flops = T * (641.8e9 + 50e9 + 5.7e9)
    pk = peaks()
    ach = flops / (ms / 1e3) / 1e12
    line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / pk["bf16_tflops_sustained"], "traffic": None, "kernel": "whole encoder (ViT GEMMs + attention + projector)",
                        "peak_source": pk["source"]}
    if args.cpu_baseline:
        from oracle import arch as A
        from oracle import restate as R
        torch.set_num_threads(os.cpu_count() or 1)
        w_cpu = {k: v.float().cpu() for k, v in sd.items()}
        px = R.preprocess_frames(host).bfloat16().float()
        with torch.no_grad():
            R.visual_embed(w_cpu, A.FULL, px[:1])
            t0 = time.perf_counter()
            ref = R.visual_embed(w_cpu, A.FULL, px)
            sec = time.perf_counter() - t0
        emb32 = vis.visual_embed(frames, normalize=True, out_dtype=torch.float32).cpu()
        line["cpu_baseline"] = {"value": T / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"all 16 frames, fp32 (the oracle of record), one batch, {sec:.1f} s"}
        line["parity"] = {"emb_maxabs": float((emb32 - ref).abs().max()), "emb_bf16_out_maxabs": float((emb.float().cpu() - ref).abs().max()),
                          "emb_bf16_floor": float((ref.bfloat16().float() - ref).abs().max()), "tolerance": 2e-2}
    print(json.dumps(line), flush=True)


def ncu_traffic(tag, chunk):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json): captured at
    40 frames per pass (M=1960, the default) and at one frame per pass (M=49); null for any other pass size."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p) and chunk in (1, 40):
        return json.load(open(p)).get(tag if chunk == 40 else tag + "_M49")
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk", type=int, default=40,
                    help="frames per decoder weight pass (1 = the reference's per-frame step; k > 1 gives identical scores and "
                         "decisions, see tests/test_gpu_loop.py::test_multi_frame_passes_equal_single_frame_steps)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-frames", type=int, default=8)
    ap.add_argument("--ref-frames", type=int, default=8)
    ap.add_argument("--keep-weights-for-cpu", action="store_true")
    ap.add_argument("--no-parity", dest="parity", action="store_false", help="skip the fp32-oracle parity block (rank 0, outside the timed regions)")
    ap.add_argument("--no-configs2", dest="configs2", action="store_false", help="skip the configs[2] frame-parallel measurement")
    ap.add_argument("--workload", default="stream", choices=["stream", "encoder16"],
                    help="stream = BASELINE configs[1] (default, the metric's configuration); encoder16 = configs[0] (encoder only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "encoder16":
        run_encoder16(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
