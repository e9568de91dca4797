"""Clock timeline of one CTA of the tcgen05 attention kernel (needs a build with MMD_NVCC_EXTRA=-DMMD_ATTN_TRACE).
Roles: 0/1 softmax group 0/1 (events: 0 wait S, 1 got S, 2 S in regs, 3 max+rescale done, 4 P computed, 5 P stored, 6 P_FULL),
2/3 MMA thread for q tile 0/1 (0 wait P, 1 got P, 2 got V, 3 PV issued, 4 wait K, 5 got K, 6 QK issued), 4/5 K/V loader
(0 wait slot, 1 got slot, 2 published), 6 misc (tile 0: 0 start, 1 setup done, 5 end; tile qi: 2 loop done, 3 O read, 4 stored)."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import _lib
lib, ctx = _lib.load(), _lib.context(0)
raw = ctypes.CDLL(_lib.LIB_PATH)
s = torch.cuda.current_stream().cuda_stream
lib.mmd_set_attention_impl(1)
buf = torch.zeros(7 * 64 * 8, dtype=torch.int32, device="cuda")
assert raw.mmd_debug_attn_trace(ctypes.c_void_p(buf.data_ptr())) == 0

def dump(name, n_tiles):
    torch.cuda.synchronize()
    b = buf.cpu().view(7, 64, 8).numpy().astype("int64") & 0xffffffff
    t0 = int(b[6, 0, 0])
    rel = lambda x: int((int(x) - t0) & 0xffffffff) if x else None
    out = {"name": name, "setup_done": rel(b[6, 0, 1]), "end": rel(b[6, 0, 5]),
           "softmax_loop_done": [rel(b[6, q, 2]) for q in range(2)], "o_read": [rel(b[6, q, 3]) for q in range(2)],
           "stored": [rel(b[6, q, 4]) for q in range(2)], "tiles": []}
    for j in range(n_tiles):
        out["tiles"].append({"j": j, "sm0": [rel(x) for x in b[0, j, :7]], "sm1": [rel(x) for x in b[1, j, :7]],
                             "mma0": [rel(x) for x in b[2, j, :7]], "mma1": [rel(x) for x in b[3, j, :7]],
                             "ldK": [rel(x) for x in b[4, j, :3]], "ldV": [rel(x) for x in b[5, j, :3]]})
    print(json.dumps(out))
    buf.zero_()
    return out

res = []
T, S, H, dh = 32, 729, 16, 72
qkv = torch.randn(T * S, 3 * H * dh, device="cuda").bfloat16()
out = torch.empty(T * S, 2 * H * dh, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    lib.mmd_vit_attention(qkv.data_ptr(), out.data_ptr(), T, S, H, dh, 1, s)
torch.cuda.synchronize(); buf.zero_()
lib.mmd_vit_attention(qkv.data_ptr(), out.data_ptr(), T, S, H, dh, 1, s)
res.append(dump("vit T=32", 28))

Hq, Hkv, dh, PAGE = 28, 4, 128, 64
n_q, L = 392, 6000
n_pages = (L + PAGE - 1) // PAGE
pool = torch.randn(n_pages, 2, Hkv, PAGE, dh, device="cuda").bfloat16()
q = torch.randn(n_q, Hq, dh, device="cuda").bfloat16()
outd = torch.empty(n_q, Hq * dh, device="cuda", dtype=torch.bfloat16)
desc = torch.tensor([0, n_q, L, 0], device="cuda", dtype=torch.int32)
tab = torch.arange(n_pages, device="cuda", dtype=torch.int32)
ns = lib.mmd_kv_attention_splits(ctx, n_q, Hq, Hkv, 1, L)
o_part = torch.empty(ns, n_q * Hq, dh, device="cuda"); ml = torch.empty(ns, n_q * Hq, 2, device="cuda")
call = lambda: lib.mmd_kv_attention(ctx, q.data_ptr(), pool.data_ptr(), desc.data_ptr(), tab.data_ptr(), 1, n_q, n_q, L,
                                    o_part.data_ptr(), ml.data_ptr(), outd.data_ptr(), Hq, Hkv, dh, ns, s)
for _ in range(2):
    _lib.check(call())
torch.cuda.synchronize(); buf.zero_()
_lib.check(call())
res.append(dump(f"kv n_q={n_q} L={L} splits={ns}", 40))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/trace_attn.json", "w"))
