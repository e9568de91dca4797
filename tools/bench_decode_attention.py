"""KV-append attention in the M = 1 decode regime (one new token over a long context): the genuinely HBM-bound regime of the
north star's '>= 70 % of HBM peak on KV-append attention' target (2048 B of K/V per context token per layer,
7 FLOP/B).  Times mmd_kv_attention alone with CUDA events over 28 per-layer pools (3.4 GB at 58.8k tokens: far beyond L2) and
reports achieved GB/s against MEASURED_PEAKS.json.  Also the 49-query frame step at the same contexts (tensor-bound).
Writes gpurun_out/decode_attention.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import _lib  # noqa: E402
from bench import ClockSampler  # noqa: E402

dev = torch.device("cuda:0")
lib, ctx = _lib.load(), _lib.context(0)
Hq, Hkv, dh, PAGE, LAYERS = 28, 4, 128, _lib.PAGE_TOKENS, 28
peak = 6552.0
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
res = []
IMPL = int(os.environ.get("MMD_ATTN_IMPL", "2"))
lib.mmd_set_attention_impl(IMPL)
for L in (5912, 29400, 58800):
    n_pages = (L + PAGE - 1) // PAGE
    pools = [torch.randn(n_pages, 2, Hkv, PAGE, dh, device=dev).bfloat16() for _ in range(LAYERS)]
    tab = torch.randperm(n_pages, device=dev).to(torch.int32)
    for n_q in (1, 49):
        q = torch.randn(n_q, Hq, dh, device=dev).bfloat16()
        desc = torch.tensor([0, n_q, L, 0], device=dev, dtype=torch.int32)
        ns = lib.mmd_kv_attention_splits(ctx, n_q, Hq, Hkv, 1, L)
        o_part = torch.empty(ns, n_q * Hq, dh, device=dev)
        ml = torch.empty(ns, n_q * Hq, 2, device=dev)
        out = torch.empty(n_q, Hq * dh, device=dev, dtype=torch.bfloat16)

        def run():
            for pool in pools:
                _lib.check(lib.mmd_kv_attention(ctx, q.data_ptr(), pool.data_ptr(), desc.data_ptr(), tab.data_ptr(), 1, n_q, n_q, L,
                                                o_part.data_ptr(), ml.data_ptr(), out.data_ptr(), Hq, Hkv, dh, ns, torch.cuda.current_stream().cuda_stream))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        # the 28 per-layer launches are replayed from a CUDA graph: a Python/ctypes call per layer (~15 us) would otherwise
        # be slower than the kernels themselves (inside mmd_decoder_step they are issued from C++ back to back)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        reps = max(10, int(600.0 / max(e0.elapsed_time(e1), 1e-3)))       # >= 0.6 s under load: the clock sampler ticks every 100 ms
        clk = ClockSampler(0)
        clk.start()
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        clocks = clk.stop()
        us = e0.elapsed_time(e1) * 1e3 / (reps * LAYERS)
        nbytes = 2 * L * Hkv * dh * 2 + 2 * n_q * Hq * dh * 2
        flops = 4.0 * n_q * L * Hq * dh
        r = {"context": L, "n_q": n_q, "splits": ns, "us_per_layer": round(us, 2), "GBps": round(nbytes / us / 1e3, 1), "hbm_frac": round(nbytes / us / 1e3 / peak, 3),
             "TFLOPs": round(flops / us / 1e6, 1), "ms_per_step_28_layers": round(us * LAYERS / 1e3, 3), "clocks": clocks}
        res.append(r)
        print(r, flush=True)
    del pools
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"hbm_peak_GBps": peak, "note": "attention kernel + split-KV combine per layer, 28 distinct pools per pass replayed from a CUDA graph, CUDA events", "results": res},
          open("gpurun_out/decode_attention.json", "w"), indent=1)
