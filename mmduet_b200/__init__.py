"""mmduet_b200 — B200-native per-frame streaming hot path of MMDuet (SigLIP -> projector/pool -> Qwen2 KV-append -> heads)."""
__version__ = "0.1.0"
