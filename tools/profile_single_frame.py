"""Stage breakdown of the live-mode encode: ONE frame through SigLIP + projector + pool (mmd_profile tags)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import _lib
from mmduet_b200.config import ModelConfig
from mmduet_b200.engine import VisionEngine
from mmduet_b200.random_init import random_state_dict, synthetic_frames
dev = torch.device("cuda:0")
cfg = ModelConfig()
sd = random_state_dict(cfg, seed=1234, device=dev, include_lm_head=False)
sd = {k: v for k, v in sd.items() if k.startswith("model.vision_tower") or k.startswith("model.mm_projector")}
vis = VisionEngine(cfg, sd, dev)
fr = synthetic_frames(4, seed=1, device=dev)
for T in (1, 4):
    for _ in range(5):
        vis.visual_embed(fr[:T], normalize=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        vis.visual_embed(fr[:T], normalize=True)
    e1.record(); torch.cuda.synchronize()
    _lib.profile_start("all", 0)
    vis.visual_embed(fr[:T], normalize=True)
    torch.cuda.synchronize()
    st = _lib.profile_stop(0)
    print(json.dumps({"frames": T, "ms_per_call": e0.elapsed_time(e1) / 20,
                      "stage_us_per_launch": {k: round(1e3 * v[0] / max(v[1], 1), 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1][0])},
                      "stage_ms": {k: round(v[0], 3) for k, v in sorted(st.items(), key=lambda kv: -kv[1][0])}}))
