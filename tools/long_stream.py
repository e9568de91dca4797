"""BASELINE configs[4]: long-stream stress — one video, 20 min @ 1 fps = 1200 frames, KV cache grown to 58.8k tokens
(beyond Qwen2's 32k max_position_embeddings; RoPE tables are built to the requested length).  Frame tokens are synthetic
(the encoder cost does not depend on the context); reports per-frame decoder latency vs context and p50/p99."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200.config import ModelConfig
from mmduet_b200.engine import DecoderEngine
from mmduet_b200.random_init import random_state_dict

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
cfg = ModelConfig()
sd = random_state_dict(cfg, seed=1234, device=dev, include_lm_head=False)
sd = {k_: v for k_, v in sd.items() if not k_.startswith("model.vision_tower") and not k_.startswith("model.mm_projector")}
ctx = 32 + n_frames * 49 + 64
dec = DecoderEngine(cfg, sd, dev, max_context=ctx, max_tokens=32 + 49 * k)
g = torch.Generator(device=dev).manual_seed(3)
bank = (torch.randn(64 * 49, cfg.hidden, device=dev, generator=g) * 1.14).bfloat16()
st, L, lat, trace = dec.new_stream(), 0, [], []
torch.cuda.synchronize()
t_all = time.perf_counter()
for f0 in range(0, n_frames, k):
    nf = min(k, n_frames - f0)
    fr = torch.cat([bank[((f0 + j) % 64) * 49:((f0 + j) % 64 + 1) * 49] for j in range(nf)], 0)
    t0 = time.perf_counter()
    out = dec.step([dict(storage=st, past=L, ids=list(range(100, 132)) if f0 == 0 else [], frames=fr,
                         score_rows=[(32 if f0 == 0 else 0) + 49 * (j + 1) - 1 for j in range(nf)])], score="frame_ends")
    sc = out["scores"].tolist()          # D2H read = sync, as the frame loop does
    dt = (time.perf_counter() - t0) * 1e3
    L = out["views"][0].length
    lat.append(dt / nf)
    if (f0 // k) % max(1, (n_frames // k) // 12) == 0:
        trace.append({"frame": f0, "context": L, "ms_per_frame": round(dt / nf, 3)})
total = time.perf_counter() - t_all
ls = sorted(lat)
rep = {"frames": n_frames, "frames_per_pass": k, "final_context": L, "kv_pool_GB": round(dec.pool.numel() * 2 / 1e9, 2),
       "p50_ms_per_frame": round(ls[len(ls) // 2], 3), "p99_ms_per_frame": round(ls[int(len(ls) * 0.99) - 1], 3),
       "decoder_frames_per_s": round(n_frames / total, 1), "last_scores": sc[-1], "trace": trace}
print(json.dumps(rep))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open(f"gpurun_out/long_stream_k{k}.json", "w"), indent=1)
