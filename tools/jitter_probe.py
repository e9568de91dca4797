"""Where are the wrong outputs of the tcgen05 attention kernel in a timing-jitter build (MMD_LIB_PATH=..._j{1,2}.so)?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import _lib
lib, ctx = _lib.load(), _lib.context(0)
s = torch.cuda.current_stream().cuda_stream
lib.mmd_set_attention_impl(1)
dh = 72
for T, S, H in [(1, 64, 1), (1, 128, 1), (1, 256, 1), (1, 729, 2), (2, 729, 16)]:
    torch.manual_seed(S)
    qkv = (torch.randn(T * S, 3 * H * dh, device="cuda") * 1.5).bfloat16()
    q, k, v = (t.view(T, S, H, dh).transpose(1, 2) for t in qkv.float().view(T, S, 3, H * dh).unbind(2))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v).transpose(1, 2).reshape(T * S, H * dh)
    worst = 0.0; nbad = 0; info = None
    for rep in range(6):
        out = torch.empty(T * S, H * dh, device="cuda", dtype=torch.bfloat16)
        _lib.check(lib.mmd_vit_attention(qkv.data_ptr(), out.data_ptr(), T, S, H, dh, 0, s))
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().view(T, S, H, dh)
        if float(err.max()) > 2e-2:
            nbad += 1
            if info is None:
                per_tile = err.amax(3)                         # [T, S, H]
                bad = (per_tile > 2e-2).nonzero()
                rows = sorted(set(int(b[1]) // 128 for b in bad)); heads = sorted(set(int(b[2]) for b in bad)); ts = sorted(set(int(b[0]) for b in bad))
                nanc = int(torch.isnan(out.float()).sum())
                info = dict(max=float(err.max()) if nanc == 0 else "nan", n_bad_rows=len(bad), q_tiles=rows, heads=heads[:8], frames=ts, nan=nanc,
                            first_bad=[bad[0].tolist(), bad[-1].tolist()])
        worst = max(worst, float(err.nan_to_num(9.0).max()))
    print((T, S, H), "bad runs", nbad, "of 6", info, flush=True)
