import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import ops, _lib
from mmduet_b200._lib import EPI_BF16, EPI_RESID_F32, ACT_GELU_TANH, ACT_NONE
lib = _lib.load()
torch.manual_seed(0)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
M = 23328
for name, N, K, epi, act in (("qkv", 3456, 1152, EPI_BF16, ACT_NONE), ("fc1", 4304, 1152, EPI_BF16, ACT_GELU_TANH),
                             ("fc2", 1152, 4304, EPI_RESID_F32, ACT_NONE), ("out_proj", 1152, 2304, EPI_RESID_F32, ACT_NONE)):
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    b = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi == EPI_RESID_F32 else torch.bfloat16)
    r = {"name": name}
    for two in (0, 1):
        lib.mmd_set_gemm_2cta(two)
        ms = timeit(lambda: ops.gemm(x, w, bias=b, act=act, out=out, epi=epi))
        r["2cta" if two else "1cta"] = {"ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
    ms = timeit(lambda: torch.matmul(x, w.t()))
    r["cublas_plain"] = {"ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
    print(json.dumps(r), flush=True)
