"""Weight-streaming GEMMs of one decoder layer in isolation: plain row-major weights vs tile-blocked weights, several
split-K settings.  Weights rotate over 6 copies so that every launch streams from HBM.  Prints GB/s of weight bytes."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import ops

dev = "cuda"
torch.manual_seed(0)
res = []

def timeit(fn, iters=30):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def case(name, M, N, K, splits_list, swiglu=False):
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    R = 6
    ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(R)]
    wb = [ops.pack_blocked(w) for w in ws]
    nbytes = N * K * 2 * (2 if swiglu else 1)
    if swiglu:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ref = ops.gemm_t_swiglu(x, ws[0], ws[1]).float()
        got = ops.gemm_t_swiglu(x, wb[0], wb[1], blocked_shape=(N, K)).float()
        err = (ref - got).abs().max().item()
        t_plain = timeit(lambda i: ops.gemm_t_swiglu(x, ws[i % R], ws[(i + 1) % R], out=out))
        t_blk = timeit(lambda i: ops.gemm_t_swiglu(x, wb[i % R], wb[(i + 1) % R], out=out, blocked_shape=(N, K)))
        r = {"name": name, "M": M, "plain_us": t_plain * 1e3, "plain_GBs": nbytes / t_plain / 1e6, "blocked_us": t_blk * 1e3,
             "blocked_GBs": nbytes / t_blk / 1e6, "maxdiff": err}
        res.append(r); print(json.dumps(r), flush=True)
        return
    for s in splits_list:
        out = ops.gemm_t_partials(x, ws[0], s)
        got = ops.gemm_t_partials(x, wb[0], s, blocked_shape=(N, K))
        err = (out.sum(0) - got.sum(0)).abs().max().item()
        t_plain = timeit(lambda i: ops.gemm_t_partials(x, ws[i % R], s, out=out))
        t_blk = timeit(lambda i: ops.gemm_t_partials(x, wb[i % R], s, out=out, blocked_shape=(N, K)))
        r = {"name": name, "M": M, "splits": s, "plain_us": t_plain * 1e3, "plain_GBs": nbytes / t_plain / 1e6,
             "blocked_us": t_blk * 1e3, "blocked_GBs": nbytes / t_blk / 1e6, "maxdiff": err}
        res.append(r); print(json.dumps(r), flush=True)

for M in (49, 196):
    case("qkv", M, 4608, 3584, [1, 2, 4, 8])
    case("o", M, 3584, 3584, [1, 2, 4, 5, 8])
    case("down", M, 3584, 18944, [2, 4, 5, 8])
    case("gate_up", M, 18944, 3584, [1], swiglu=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/decoder_gemms.json", "w"), indent=1)
