"""Weight ingestion helpers: the reference loads `lmms-lab/llava-onevision-qwen2-7b-ov` and an
UNMERGED peft LoRA adapter (models/modeling_live.py:96-123; r=16, alpha=32 on q/k/v/o/gate/up/down of every decoder layer,
models/arguments_live.py:12-15).  The kernels consume plain weight matrices, so the adapter is merged once at load:
W' = W + (alpha / r) * B @ A  (peft's own merge formula), computed in fp32 and rounded to bf16 once."""
import re

import torch

_LORA_RE = re.compile(r"^(?:base_model\.model\.)?(?P<mod>.+?)\.lora_(?P<ab>[AB])(?:\.default)?\.weight$")
_BASE_RE = re.compile(r"^(?:base_model\.model\.)?(?P<mod>.+?)(?:\.base_layer)?\.(?P<leaf>weight|bias)$")


def merge_lora(state_dict, lora_state_dict=None, lora_r=16, lora_alpha=32):
    """Returns a new state_dict with every `*.lora_A/B` pair folded into its base weight.

    Accepts either a separate adapter state_dict (keys as saved by peft: `base_model.model.<module>.lora_A.weight`) or a
    single dict that already contains base (`<module>.base_layer.weight`) and adapter tensors.  Keys of the result are
    the reference's plain module names (`model.layers.N.self_attn.q_proj.weight`, ...)."""
    pairs, out = {}, {}
    sources = [state_dict] + ([lora_state_dict] if lora_state_dict else [])
    for sd in sources:
        for k, v in sd.items():
            m = _LORA_RE.match(k)
            if m:
                pairs.setdefault(m.group("mod"), {})[m.group("ab")] = v
                continue
            if sd is lora_state_dict and "lora_" not in k and k.startswith("base_model.model.") is False and k in out:
                continue
            m = _BASE_RE.match(k)
            out[f"{m.group('mod')}.{m.group('leaf')}" if m else k] = v
    scale = float(lora_alpha) / float(lora_r)
    for mod, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise ValueError(f"incomplete LoRA pair for {mod}")
        key = mod + ".weight"
        if key not in out:
            raise KeyError(f"LoRA adapter targets {mod} but the base state_dict has no {key}")
        w = out[key]
        delta = ab["B"].to(torch.float32) @ ab["A"].to(torch.float32)
        if delta.shape != w.shape:
            raise ValueError(f"LoRA shape mismatch for {mod}: {tuple(delta.shape)} vs {tuple(w.shape)}")
        out[key] = (w.to(torch.float32) + scale * delta.to(w.device)).to(w.dtype)
    return out
