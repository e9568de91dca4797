"""Architecture/config of the accelerated path (mirrors the fields the reference reads from its HF configs:
models/arguments_live.py:19-22, video_head_live_llava_qwen.py:41-45,100-119; SigLIP-so400m/14@384 + Qwen2-7B sizes)."""
from dataclasses import dataclass, fields


@dataclass(frozen=True)
class ModelConfig:
    # SigLIP vision tower
    image_size: int = 384
    patch_size: int = 14
    vit_dim: int = 1152
    vit_heads: int = 16
    vit_mlp: int = 4304
    vit_layers_total: int = 27
    # Qwen2 decoder
    hidden: int = 3584
    layers: int = 28
    q_heads: int = 28
    kv_heads: int = 4
    mlp: int = 18944
    vocab: int = 152064
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    max_pos: int = 32768
    # pooling
    pool_stride: int = 4
    pool_mode: str = "bilinear"
    frame_tokens: int = 49

    @property
    def vit_layers(self):
        """Layers executed on the llava path (LLaVA's SigLipVisionTower deletes the last encoder layer)."""
        return self.vit_layers_total - 1

    @property
    def grid(self):
        return self.image_size // self.patch_size

    @property
    def patches(self):
        return self.grid * self.grid

    @property
    def head_dim(self):
        return self.hidden // self.q_heads

    @property
    def vit_head_dim(self):
        return self.vit_dim // self.vit_heads

    @classmethod
    def from_any(cls, obj):
        """Builds a ModelConfig from any object/dict exposing the same field names (e.g. a test-side arch record)."""
        get = (lambda k: obj[k]) if isinstance(obj, dict) else (lambda k: getattr(obj, k))
        return cls(**{f.name: get(f.name) for f in fields(cls)})

    def validate(self):
        if self.vit_head_dim != 72:
            raise ValueError(f"vision head_dim {self.vit_head_dim} unsupported: the fused attention kernel is built for 72")
        if self.head_dim != 128:
            raise ValueError(f"decoder head_dim {self.head_dim} unsupported: the KV-append attention kernel is built for 128")
        for name in ("vit_dim", "vit_mlp", "hidden", "mlp"):
            if getattr(self, name) % 8:
                raise ValueError(f"{name} must be a multiple of 8 (16-byte TMA rows)")
        if self.q_heads % self.kv_heads:
            raise ValueError("q_heads must be a multiple of kv_heads")
        if self.pool_mode not in ("bilinear", "average", "max"):
            raise ValueError(f"Unexpected mm_spatial_pool_mode: {self.pool_mode}")
