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
