import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import ops, _lib
from mmduet_b200._lib import EPI_BF16, EPI_RESID_F32, EPI_F32, ACT_NONE
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
for name, N, K in (("fc2", 1152, 4304), ("out_proj", 1152, 2304), ("N1024", 1024, 4304), ("N1280", 1280, 4304)):
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    b = torch.randn(N, device="cuda")
    r = {"name": name, "bn192": os.environ.get("MMD_NO_BN192") is None}
    for epi, en in ((EPI_RESID_F32, "resid"), (EPI_F32, "f32"), (EPI_BF16, "bf16")):
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if epi == EPI_BF16 else torch.float32)
        ms = timeit(lambda: ops.gemm(x, w, bias=b, out=out, epi=epi))
        r[en] = round(2.0 * M * N * K / ms / 1e9, 1)
    print(json.dumps(r), flush=True)
