"""Target for `ncu --set full -k regex:kv_decode_attention`: a few M = 1 decode-attention launches over 58.8k-token contexts
(28 distinct per-layer pools, 3.4 GB: nothing is L2-resident), no CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
lib, ctx = _lib.load(), _lib.context(0)
Hq, Hkv, dh, PAGE, L = 28, 4, 128, _lib.PAGE_TOKENS, int(os.environ.get("CTX", 58800))
n_pages = (L + PAGE - 1) // PAGE
pools = [torch.randn(n_pages, 2, Hkv, PAGE, dh, device=dev).bfloat16() for _ in range(28)]
tab = torch.randperm(n_pages, device=dev).to(torch.int32)
q = torch.randn(1, Hq, dh, device=dev).bfloat16()
desc = torch.tensor([0, 1, L, 0], device=dev, dtype=torch.int32)
out = torch.empty(1, Hq * dh, device=dev, dtype=torch.bfloat16)
ns = lib.mmd_kv_attention_splits(ctx, 1, Hq, Hkv, 1, L)
o_part = torch.empty(max(ns, 64) * 2, Hq, dh, device=dev)
ml = torch.empty(max(ns, 64) * 2, Hq, 2, device=dev)
s = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    for pool in pools:
        _lib.check(lib.mmd_kv_attention(ctx, q.data_ptr(), pool.data_ptr(), desc.data_ptr(), tab.data_ptr(), 1, 1, 1, L, o_part.data_ptr(),
                                        ml.data_ptr(), out.data_ptr(), Hq, Hkv, dh, ns, s))
torch.cuda.synchronize()
print("algorithmic bytes per launch", L * Hkv * dh * 2 * 2, "splits", ns)
