"""mmduet_b200 — B200-native per-frame streaming hot path of MMDuet (SigLIP -> projector/pool -> Qwen2 KV-append -> heads).

Public surface mirrors the reference's (models/__init__.py:8-20, test/inference.py, demo/liveinfer.py)."""
__version__ = "0.1.0"

from .config import ModelConfig  # noqa: F401


def build_model_and_tokenizer(is_training=False, *, state_dict=None, model_config=None, device="cuda", tokenizer=None,
                              max_context=None, kv_pages=None, **kwargs):
    """models/__init__.py:8-13.  Checkpoints cannot be downloaded offline, so the weights come from `state_dict` (keys as
    in the reference's checkpoint: model.vision_tower..., model.mm_projector..., model.layers..., lm_head,
    informative_head, relevance_head; LoRA deltas must be merged by the caller: W + alpha/r * B @ A)."""
    if is_training:
        raise NotImplementedError("training is outside the accelerated path")
    if state_dict is None:
        raise ValueError("build_model_and_tokenizer needs state_dict= (no network access to fetch "
                         f"{kwargs.get('llm_pretrained', 'the checkpoint')})")
    from .modeling_live import VideoHeadLiveLlavaQwenForCausalLM
    from .tokenization_live import SyntheticTokenizer
    cfg = model_config or ModelConfig()
    if tokenizer is None:
        tokenizer = SyntheticTokenizer(cfg.vocab)
    # what build_live_tokenizer_and_update_config writes into the model config (models/tokenization_live.py:118-124)
    v_id = tokenizer.convert_tokens_to_ids("<image>") if hasattr(tokenizer, "convert_tokens_to_ids") else None
    model = VideoHeadLiveLlavaQwenForCausalLM(cfg, state_dict, device=device, max_context=max_context, kv_pages=kv_pages,
                                              eos_token_id=getattr(tokenizer, "eos_token_id", None), v_placeholder_id=v_id)
    return model, tokenizer
