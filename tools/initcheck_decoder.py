"""Tiny decoder scenario for compute-sanitizer (initcheck / memcheck): frame steps, two generated tokens, rollback, another
frame step — TINY architecture, no ViT."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200.config import ModelConfig
from mmduet_b200.engine import DecoderEngine
from oracle import arch as A, restate as R

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
arch = A.SMALL
w = R.make_weights(arch, seed=11)
cfg = ModelConfig.from_any(arch)
dec = DecoderEngine(cfg, w, dev, max_context=1024, max_lm_rows=1)
g = torch.Generator().manual_seed(1)
fr = (torch.randn(4 * 49, cfg.hidden, generator=g) * 1.1).bfloat16().to(dev)
tpf = 49
long_st, LL = dec.new_stream(), 0
for f in range(12):     # a longer stream first: more key tiles, split-KV, page boundaries
    o_ = dec.step([dict(storage=long_st, past=LL, ids=[5, 6, 7, 8, 9, 10] if f == 0 else [], frames=fr[(f % 4) * tpf:(f % 4 + 1) * tpf])])
    LL = o_["views"][0].length
print("long stream", LL, o_["scores"].tolist())
st, L = dec.new_stream(), 0
out = dec.step([dict(storage=st, past=L, ids=[5, 6, 7, 8], frames=fr[:49])]); L = out["views"][0].length
out = dec.step([dict(storage=st, past=L, ids=[], frames=fr[49:98])]); L = out["views"][0].length
keep = L
emb = torch.randn(3, cfg.hidden, generator=g).bfloat16().to(dev)
o = dec.step([dict(storage=st, past=L, embeds=emb)], score="none", lm="last"); L = o["views"][0].length
for _ in range(2):
    tok = int(o["lm_logits"].argmax())
    e = dec.embed[tok:tok + 1]
    o = dec.step([dict(storage=st, past=L, embeds=e)], score="none", lm="last"); L = o["views"][0].length
out2 = dec.step([dict(storage=st, past=keep, ids=[], frames=fr[98:147])])      # rollback to `keep`, next frame
torch.cuda.synchronize()
print("scores", out["scores"].tolist(), out2["scores"].tolist())
