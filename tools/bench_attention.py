"""mma.sync vs tcgen05 attention kernels in isolation (ViT batch 32; decoder KV-append at several chunk sizes/contexts)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import _lib
lib, ctx = _lib.load(), _lib.context(0)
s = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)

from bench import ClockSampler
LAST_CLOCKS = {}

def timeit(fn, iters=20):
    """us per call; the timed loop runs >= 0.5 s so that nvidia-smi (100 ms ticks) samples the clocks under this load"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    iters = max(iters, int(500.0 / max(e0.elapsed_time(e1) / iters, 1e-3)))
    clk = ClockSampler(0); clk.start()
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    LAST_CLOCKS.clear(); LAST_CLOCKS.update(clk.stop())
    return e0.elapsed_time(e1) / iters * 1e3

res = []
T, S, H, dh = 32, 729, 16, 72
qkv = (torch.randn(T * S, 3 * H * dh, device="cuda")).bfloat16()
out = torch.empty(T * S, 2 * H * dh, device="cuda", dtype=torch.bfloat16)
flops = 4.0 * T * H * S * S * dh
for impl in (0, 1):
    lib.mmd_set_attention_impl(impl)
    us = timeit(lambda: lib.mmd_vit_attention(qkv.data_ptr(), out.data_ptr(), T, S, H, dh, 1, s()))
    res.append({"kernel": "vit_attention", "impl": impl, "us": us, "tflops": flops / us / 1e6, "clocks": dict(LAST_CLOCKS)}); print(res[-1], flush=True)

Hq, Hkv, dh, PAGE = 28, 4, 128, 64
for n_q, L in ((49, 3000), (49, 30000), (392, 3000), (392, 6000), (1, 6000)):
    n_pages = (L + PAGE - 1) // PAGE
    pool = torch.randn(n_pages, 2, Hkv, PAGE, dh, device="cuda").bfloat16()
    q = torch.randn(n_q, Hq, dh, device="cuda").bfloat16()
    desc = torch.tensor([0, n_q, L, 0], device="cuda", dtype=torch.int32)
    tab = torch.arange(n_pages, device="cuda", dtype=torch.int32)
    outd = torch.empty(n_q, Hq * dh, device="cuda", dtype=torch.bfloat16)
    fl = 4.0 * n_q * (L - n_q / 2) * Hq * dh
    for impl in (0, 1):
        lib.mmd_set_attention_impl(impl)
        ns = lib.mmd_kv_attention_splits(ctx, n_q, Hq, Hkv, 1, L)
        o_part = torch.empty(ns, n_q * Hq, dh, device="cuda"); ml = torch.empty(ns, n_q * Hq, 2, device="cuda")
        us = timeit(lambda: lib.mmd_kv_attention(ctx, q.data_ptr(), pool.data_ptr(), desc.data_ptr(), tab.data_ptr(), 1, n_q, n_q, L,
                                                 o_part.data_ptr(), ml.data_ptr(), outd.data_ptr(), Hq, Hkv, dh, ns, s()))
        res.append({"kernel": "kv_attention", "n_q": n_q, "L": L, "impl": impl, "splits": ns, "us": us, "tflops": fl / us / 1e6,
                    "kv_GBs": L * Hkv * dh * 2 * 2 / us / 1e3, "clocks": dict(LAST_CLOCKS)}); print(res[-1], flush=True)
lib.mmd_set_attention_impl(2)
json.dump(res, open("gpurun_out/bench_attention.json", "w"), indent=1)
