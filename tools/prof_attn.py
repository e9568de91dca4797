import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import _lib
lib, ctx = _lib.load(), _lib.context(0)
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib.mmd_set_attention_impl(impl)
T, S, H, dh = 32, 729, 16, 72
qkv = torch.randn(T * S, 3 * H * dh, device="cuda").bfloat16()
out = torch.empty(T * S, 2 * H * dh, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    lib.mmd_vit_attention(qkv.data_ptr(), out.data_ptr(), T, S, H, dh, 1, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
# decoder KV-append attention: 10 frames (490 query tokens) appended at 3k context, one layer
Hq, Hkv, dh, PAGE = 28, 4, 128, 64
n_q, L = 490, 3490
n_pages = (L + PAGE - 1) // PAGE
pool = torch.randn(n_pages, 2, Hkv, PAGE, dh, device="cuda").bfloat16()
q = torch.randn(n_q, Hq, dh, device="cuda").bfloat16()
desc = torch.tensor([0, n_q, L, 0], device="cuda", dtype=torch.int32)
tab = torch.arange(n_pages, device="cuda", dtype=torch.int32)
outd = torch.empty(n_q, Hq * dh, device="cuda", dtype=torch.bfloat16)
ns = lib.mmd_kv_attention_splits(ctx, n_q, Hq, Hkv, 1, L)
o_part = torch.empty(ns, n_q * Hq, dh, device="cuda"); ml = torch.empty(ns, n_q * Hq, 2, device="cuda")
for _ in range(3):
    lib.mmd_kv_attention(ctx, q.data_ptr(), pool.data_ptr(), desc.data_ptr(), tab.data_ptr(), 1, n_q, n_q, L,
                         o_part.data_ptr(), ml.data_ptr(), outd.data_ptr(), Hq, Hkv, dh, ns, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
