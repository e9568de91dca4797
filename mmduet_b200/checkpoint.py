"""Weight ingestion.  The reference loads `lmms-lab/llava-onevision-qwen2-7b-ov` with
`from_pretrained` and wraps it in an UNMERGED peft LoRA adapter (models/modeling_live.py:96-123; r=16, alpha=32 on
q/k/v/o/gate/up/down of every decoder layer, mm_projector and the two heads saved as full `modules_to_save` copies,
models/arguments_live.py:12-15).  Here:

  * `load_safetensors(path)` reads a `.safetensors` file, a sharded checkpoint directory (`model.safetensors.index.json`) or
    a peft adapter directory (`adapter_model.safetensors` + `adapter_config.json`) into a flat state_dict;
  * `merge_lora(base, adapter)` folds every `lora_A/lora_B` pair into its base matrix, W' = W + (alpha / r) * B @ A (peft's own
    merge formula) in fp32, rounded to bf16 once, and lets `modules_to_save` copies replace their base modules — the kernels
    consume plain matrices, and tests/test_gpu_parity.py::test_merged_lora_matches_unmerged_forward holds the merged model to
    the north-star tolerance against the unmerged fp32 forward;
  * `state_dict_from_pretrained(llm_pretrained, lora_pretrained)` is what `build_model_and_tokenizer` calls for local paths."""
import json
import os
import re

import torch

_LORA_RE = re.compile(r"^(?:base_model\.model\.)?(?P<mod>.+?)\.lora_(?P<ab>[AB])(?:\.[A-Za-z0-9_]+)?\.weight$")
_SAVE_RE = re.compile(r"^(?:base_model\.model\.)?(?P<mod>.+?)\.modules_to_save(?:\.[A-Za-z0-9_]+)?\.(?P<leaf>weight|bias)$")
_ORIG_RE = re.compile(r"^(?:base_model\.model\.)?(?P<mod>.+?)\.original_module\.(?P<leaf>weight|bias)$")
_BASE_RE = re.compile(r"^(?:base_model\.model\.)?(?P<mod>.+?)(?:\.base_layer)?\.(?P<leaf>weight|bias)$")


def load_safetensors(path, device="cpu"):
    """-> (state_dict, meta).  `path`: a .safetensors file, or a directory holding model.safetensors[.index.json] shards or a
    peft adapter (adapter_model.safetensors [+ adapter_config.json, returned as meta])."""
    from safetensors import safe_open
    files, meta = [], {}
    if os.path.isdir(path):
        idx = os.path.join(path, "model.safetensors.index.json")
        if os.path.exists(idx):
            files = sorted({os.path.join(path, f) for f in json.load(open(idx))["weight_map"].values()})
        else:
            for name in ("model.safetensors", "adapter_model.safetensors"):
                if os.path.exists(os.path.join(path, name)):
                    files.append(os.path.join(path, name))
        cfg = os.path.join(path, "adapter_config.json")
        if os.path.exists(cfg):
            meta = json.load(open(cfg))
    elif os.path.exists(path):
        files = [path]
    if not files:
        raise FileNotFoundError(f"no safetensors checkpoint at {path} (there is no network access to fetch a hub id)")
    sd = {}
    for f in files:
        with safe_open(f, framework="pt", device=str(device)) as fh:
            for k in fh.keys():
                sd[k] = fh.get_tensor(k)
    return sd, meta


def merge_lora(state_dict, lora_state_dict=None, lora_r=16, lora_alpha=32):
    """Returns a new state_dict with plain module names (`model.layers.N.self_attn.q_proj.weight`, ...): every `*.lora_A/B`
    pair folded into its base weight, every `modules_to_save` copy in place of its base module.

    Accepts a separate adapter state_dict as saved by peft (`base_model.model.<module>.lora_A.weight`, full copies of the
    modules_to_save under their plain names) and/or a live PeftModel.state_dict() (`<module>.base_layer.weight`,
    `<module>.lora_A.default.weight`, `<module>.modules_to_save.default.weight`, `<module>.original_module.weight`)."""
    pairs, saved, out = {}, {}, {}
    for sd in [state_dict] + ([lora_state_dict] if lora_state_dict else []):
        is_adapter = sd is lora_state_dict
        for k, v in sd.items():
            m = _LORA_RE.match(k)
            if m:
                pairs.setdefault(m.group("mod"), {})[m.group("ab")] = v
                continue
            m = _SAVE_RE.match(k)
            if m:
                saved[f"{m.group('mod')}.{m.group('leaf')}"] = v
                continue
            if _ORIG_RE.match(k):
                continue                                  # superseded by the modules_to_save copy
            m = _BASE_RE.match(k)
            name = f"{m.group('mod')}.{m.group('leaf')}" if m else k
            if is_adapter:
                saved[name] = v                           # a saved adapter lists its modules_to_save under their plain names
            else:
                out[name] = v
    for name, v in saved.items():
        if name in out and out[name].shape != v.shape:
            raise ValueError(f"modules_to_save copy of {name} has shape {tuple(v.shape)}, base has {tuple(out[name].shape)}")
        out[name] = v.to(out[name].dtype) if name in out else v
    scale = float(lora_alpha) / float(lora_r)
    for mod, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise ValueError(f"incomplete LoRA pair for {mod}")
        key = mod + ".weight"
        if key not in out:
            raise KeyError(f"LoRA adapter targets {mod} but the base state_dict has no {key}")
        w = out[key]
        if ab["A"].shape[0] != lora_r and lora_state_dict is not None and ab["A"].shape[0] != ab["B"].shape[1]:
            raise ValueError(f"LoRA rank mismatch for {mod}")
        delta = ab["B"].to(device=w.device, dtype=torch.float32) @ ab["A"].to(device=w.device, dtype=torch.float32)
        if delta.shape != w.shape:
            raise ValueError(f"LoRA shape mismatch for {mod}: {tuple(delta.shape)} vs {tuple(w.shape)}")
        out[key] = (w.to(torch.float32) + scale * delta).to(w.dtype)
    return out


def state_dict_from_pretrained(llm_pretrained, lora_pretrained=None, lora_r=None, lora_alpha=None, device="cpu"):
    """Local-path version of the reference's from_pretrained + PeftModel.from_pretrained (models/modeling_live.py:96-123)."""
    sd, _ = load_safetensors(llm_pretrained, device)
    if lora_pretrained:
        adapter, meta = load_safetensors(lora_pretrained, device)
        sd = merge_lora(sd, adapter, lora_r=lora_r or meta.get("r", 16), lora_alpha=lora_alpha or meta.get("lora_alpha", 32))
    return sd
