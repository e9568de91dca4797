"""mmduet_b200 — B200-native per-frame streaming hot path of MMDuet (SigLIP -> projector/pool -> Qwen2 KV-append -> heads).

Public surface mirrors the reference's (models/__init__.py:8-20, test/inference.py, demo/liveinfer.py)."""
__version__ = "0.1.0"

from .config import ModelConfig  # noqa: F401


def build_model_and_tokenizer(is_training=False, *, state_dict=None, lora_state_dict=None, model_config=None, device="cuda",
                              tokenizer=None, max_context=None, kv_pages=None, **kwargs):
    """models/__init__.py:8-13 / models/modeling_live.py:80-123.  Weights come either from `state_dict` (keys as in the reference's
    checkpoint: model.vision_tower..., model.mm_projector..., model.layers..., lm_head, informative_head, relevance_head) or
    from `llm_pretrained` when that is a LOCAL safetensors checkpoint directory (hub ids cannot be fetched: no network).
    A LoRA adapter (`lora_state_dict`, or `lora_pretrained` as a local peft adapter directory; r / alpha from its
    adapter_config.json or the `lora_r` / `lora_alpha` flags) is merged into the base matrices once at load.  The tokenizer is
    `tokenizer`, else the one in the `llm_pretrained` directory with the live chat template, else the synthetic stand-in."""
    import os
    if is_training:
        raise NotImplementedError("training is outside the accelerated path")
    from . import checkpoint
    from .modeling_live import VideoHeadLiveLlavaQwenForCausalLM
    from .tokenization_live import SyntheticTokenizer, build_live_tokenizer_and_update_config
    llm, lora = kwargs.get("llm_pretrained"), kwargs.get("lora_pretrained")
    r, alpha = kwargs.get("lora_r") or 16, kwargs.get("lora_alpha") or 32
    if state_dict is None:
        if not (llm and os.path.exists(llm)):
            raise ValueError(f"build_model_and_tokenizer needs state_dict= or a local llm_pretrained checkpoint (no network access to fetch {llm})")
        state_dict = checkpoint.state_dict_from_pretrained(llm, lora if lora and os.path.exists(lora) else None, kwargs.get("lora_r"),
                                                           kwargs.get("lora_alpha"))
    elif lora_state_dict is not None:
        state_dict = checkpoint.merge_lora(state_dict, lora_state_dict, lora_r=r, lora_alpha=alpha)
    elif lora and os.path.exists(lora):
        adapter, meta = checkpoint.load_safetensors(lora)
        state_dict = checkpoint.merge_lora(state_dict, adapter, lora_r=kwargs.get("lora_r") or meta.get("r", 16),
                                           lora_alpha=kwargs.get("lora_alpha") or meta.get("lora_alpha", 32))
    cfg = model_config or ModelConfig()
    if tokenizer is None and llm and os.path.isdir(llm) and any(os.path.exists(os.path.join(llm, f)) for f in ("tokenizer.json", "vocab.json")):
        tokenizer = build_live_tokenizer_and_update_config(llm, dict(v_placeholder=kwargs.get("v_placeholder", "<image>"),
                                                                     frame_num_tokens=cfg.frame_tokens))
    if tokenizer is None:
        tokenizer = SyntheticTokenizer(cfg.vocab)
    # what build_live_tokenizer_and_update_config writes into the model config (models/tokenization_live.py:118-124)
    v_id = tokenizer.convert_tokens_to_ids("<image>") if hasattr(tokenizer, "convert_tokens_to_ids") else None
    model = VideoHeadLiveLlavaQwenForCausalLM(cfg, state_dict, device=device, max_context=max_context, kv_pages=kv_pages,
                                              eos_token_id=getattr(tokenizer, "eos_token_id", None), v_placeholder_id=v_id)
    return model, tokenizer
