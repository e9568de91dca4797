"""Run-to-run determinism of individual kernels (used under compute-sanitizer memcheck, where a rare mismatch showed up)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import _lib, ops
lib, ctx = _lib.load(), _lib.context(0)
s = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
REP = int(os.environ.get("REP", "12"))

def check(name, fn):
    ref = fn().clone(); bad = []
    for i in range(REP):
        o = fn()
        if not torch.equal(o, ref):
            d = (o.float() - ref.float()).abs()
            bad.append((i, float(d.max()), int((d > 0).sum())))
    print(name, "mismatches", bad if bad else "none", flush=True)

T, S, H, dh = 14, 729, 4, 72
qkv = (torch.randn(T * S, 3 * H * dh, device="cuda") * 1.5).bfloat16()
out = torch.empty(T * S, 2 * H * dh, device="cuda", dtype=torch.bfloat16)
for impl in (1, 0):
    lib.mmd_set_attention_impl(impl)
    def attn():
        out.zero_()
        _lib.check(lib.mmd_vit_attention(qkv.data_ptr(), out.data_ptr(), T, S, H, dh, 1, s()))
        return out
    check(f"vit_attention impl={impl} (168 items)", attn)
lib.mmd_set_attention_impl(2)
M, N, K = T * S, 288, 1152
x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); wgt = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
check("gemm bf16 (CTA pair)", lambda: ops.gemm(x, wgt))
from mmduet_b200._lib import EPI_RESID_F32
res0 = torch.randn(M, N, device="cuda")
def resid():
    r = res0.clone()
    ops.gemm(x, wgt, out=r, epi=EPI_RESID_F32)
    return r
check("gemm resid (TMA reduce-add)", resid)
xf = torch.randn(M, 288, device="cuda"); g = torch.randn(288, device="cuda"); b = torch.randn(288, device="cuda")
o2 = torch.empty(M, 288, device="cuda", dtype=torch.bfloat16)
def ln():
    _lib.check(lib.mmd_layernorm(xf.data_ptr(), g.data_ptr(), b.data_ptr(), o2.data_ptr(), 0, M, 288, 1e-6, s()))
    return o2
check("layernorm", ln)

# the chain that showed the memcheck-only mismatch: CTA-pair GEMM (TMA store) -> tcgen05 attention -> CTA-pair GEMM (reduce-add)
D = 288
xin = (torch.randn(M, D, device="cuda") * 0.5).bfloat16()
wqkv = (torch.randn(3 * D, D, device="cuda") * 0.06).bfloat16(); wo = (torch.randn(D, 2 * D, device="cuda") * 0.05).bfloat16()
qkv2 = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
att2 = torch.empty(M, 2 * D, device="cuda", dtype=torch.bfloat16)
def chain():
    ops.gemm(xin, wqkv, out=qkv2)
    _lib.check(lib.mmd_vit_attention(qkv2.data_ptr(), att2.data_ptr(), T, S, H, dh, 1, s()))
    r = res0.clone()
    ops.gemm(att2, wo, out=r, epi=EPI_RESID_F32)
    return r
check("chain qkv-gemm -> attention -> out_proj", chain)
