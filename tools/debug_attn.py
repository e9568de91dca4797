"""Error statistics of the tcgen05 attention kernels against fp32 torch (debug aid)."""
import sys, torch
sys.path.insert(0, ".")
from mmduet_b200 import _lib
lib, ctx = _lib.load(), _lib.context(0)
s = torch.cuda.current_stream().cuda_stream
dh = 72
for impl in (0, 1):
    lib.mmd_set_attention_impl(impl)
    for T, S, H in [(1, 64, 1), (1, 128, 1), (1, 129, 1), (1, 256, 2), (2, 729, 16)]:
        torch.manual_seed(S)
        qkv = (torch.randn(T * S, 3 * H * dh, device="cuda") * 1.5).bfloat16()
        out = torch.empty(T * S, H * dh, device="cuda", dtype=torch.bfloat16)
        _lib.check(lib.mmd_vit_attention(qkv.data_ptr(), out.data_ptr(), T, S, H, dh, 0, s))
        torch.cuda.synchronize()
        q, k, v = (t.view(T, S, H, dh).transpose(1, 2) for t in qkv.float().view(T, S, 3, H * dh).unbind(2))
        att = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1)
        ref = (att @ v).transpose(1, 2).reshape(T * S, H * dh)
        err = (out.float() - ref).abs()
        bad = (err > 2e-2).nonzero()
        print(impl, (T, S, H), "max", float(err.max()), "mean", float(err.mean()), "nbad", len(bad), "first bad", bad[:4].tolist(),
              "nan", int(torch.isnan(out.float()).sum()))
