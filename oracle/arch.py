"""TEST INFRASTRUCTURE ONLY — architecture constants shared by the oracle, the golden-vector generator and the tests.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import oracle/.
Values: SURVEY.md §8 (SigLIP-so400m/14@384 + Qwen2-7B; recalled, checkpoint config not available offline)."""
from dataclasses import dataclass, replace


@dataclass(frozen=True)
class Arch:
    # vision tower (SigLIP)
    image_size: int = 384
    patch_size: int = 14
    vit_dim: int = 1152
    vit_heads: int = 16
    vit_mlp: int = 4304
    vit_layers_total: int = 27      # HF checkpoint depth; the llava path deletes the last one
    # decoder (Qwen2)
    hidden: int = 3584
    layers: int = 28
    q_heads: int = 28
    kv_heads: int = 4
    mlp: int = 18944
    vocab: int = 152064
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    max_pos: int = 32768
    # frame pooling (arguments_live.py:20-21; video_head_live_llava_qwen.py:100-119)
    pool_stride: int = 4
    pool_mode: str = "bilinear"
    frame_tokens: int = 49

    @property
    def vit_layers(self):           # layers executed on the llava path
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


FULL = Arch()

# Same code paths, sizes the CPU oracle finishes in seconds.  Keeps the awkward properties of the real model:
# ViT head_dim 72 (not a multiple of 16), MLP width not a multiple of 128, GQA group 7, grid 27 -> 7 bilinear.
SMALL = Arch(vit_dim=288, vit_heads=4, vit_mlp=1000, vit_layers_total=4,
             hidden=896, layers=3, q_heads=7, kv_heads=1, mlp=2432, vocab=2048)

# Tiny: for fixtures committed to git (weights regenerated from a seed, outputs stored).
TINY = Arch(vit_dim=144, vit_heads=2, vit_mlp=328, vit_layers_total=3,
            hidden=256, layers=2, q_heads=2, kv_heads=1, mlp=512, vocab=512)


def with_pool(arch, mode):
    return replace(arch, pool_mode=mode)
