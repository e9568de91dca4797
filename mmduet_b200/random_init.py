"""Seeded random-init weights of the named architecture, generated on the device (there is no network for checkpoints):
state_dict with the reference's key names, bf16 for matrices / fp32 for norms and biases.  Scales follow the HF / torch
default initialisers (SigLIP lecun/xavier, nn.Linear default for mm_projector, normal(0, 0.02) for Qwen2)."""
import torch

from .config import ModelConfig

VT = "model.vision_tower.vision_tower.vision_model."


def random_state_dict(cfg: ModelConfig, seed=1234, device="cuda", include_lm_head=True, legacy_post_ln=False):
    g = torch.Generator(device=device).manual_seed(seed)
    w = {}

    def mat(*shape, std):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std).to(torch.bfloat16)

    def vec(n, std, base=0.0):
        return (base + torch.randn(n, generator=g, device=device, dtype=torch.float32) * std).bfloat16().float()

    D, Dm, P = cfg.vit_dim, cfg.vit_mlp, cfg.patch_size
    w[VT + "embeddings.patch_embedding.weight"] = mat(D, 3, P, P, std=(3 * P * P) ** -0.5)
    w[VT + "embeddings.patch_embedding.bias"] = vec(D, 0.02)
    w[VT + "embeddings.position_embedding.weight"] = mat(cfg.patches, D, std=D ** -0.5)
    for i in range(cfg.vit_layers_total if legacy_post_ln else cfg.vit_layers):
        p = f"{VT}encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            w[p + ln + ".weight"] = vec(D, 0.05, 1.0)
            w[p + ln + ".bias"] = vec(D, 0.05)
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[p + f"self_attn.{proj}.weight"] = mat(D, D, std=D ** -0.5)
            w[p + f"self_attn.{proj}.bias"] = vec(D, 0.02)
        w[p + "mlp.fc1.weight"] = mat(Dm, D, std=(2.0 / (D + Dm)) ** 0.5)
        w[p + "mlp.fc1.bias"] = vec(Dm, 0.02)
        w[p + "mlp.fc2.weight"] = mat(D, Dm, std=(2.0 / (D + Dm)) ** 0.5)
        w[p + "mlp.fc2.bias"] = vec(D, 0.02)
    if legacy_post_ln:
        w[VT + "post_layernorm.weight"] = vec(D, 0.05, 1.0)
        w[VT + "post_layernorm.bias"] = vec(D, 0.05)
    H, kvd = cfg.hidden, cfg.kv_heads * cfg.head_dim
    w["model.mm_projector.0.weight"] = mat(H, D, std=(3 * D) ** -0.5)
    w["model.mm_projector.0.bias"] = vec(H, 0.02)
    w["model.mm_projector.2.weight"] = mat(H, H, std=(3 * H) ** -0.5)
    w["model.mm_projector.2.bias"] = vec(H, 0.02)
    w["model.embed_tokens.weight"] = mat(cfg.vocab, H, std=0.02)
    for i in range(cfg.layers):
        p = f"model.layers.{i}."
        w[p + "input_layernorm.weight"] = vec(H, 0.05, 1.0)
        w[p + "post_attention_layernorm.weight"] = vec(H, 0.05, 1.0)
        w[p + "self_attn.q_proj.weight"] = mat(H, H, std=0.02)
        w[p + "self_attn.q_proj.bias"] = vec(H, 0.1)
        w[p + "self_attn.k_proj.weight"] = mat(kvd, H, std=0.02)
        w[p + "self_attn.k_proj.bias"] = vec(kvd, 0.1)
        w[p + "self_attn.v_proj.weight"] = mat(kvd, H, std=0.02)
        w[p + "self_attn.v_proj.bias"] = vec(kvd, 0.1)
        w[p + "self_attn.o_proj.weight"] = mat(H, H, std=0.02)
        w[p + "mlp.gate_proj.weight"] = mat(cfg.mlp, H, std=0.02)
        w[p + "mlp.up_proj.weight"] = mat(cfg.mlp, H, std=0.02)
        w[p + "mlp.down_proj.weight"] = mat(H, cfg.mlp, std=0.02)
    w["model.norm.weight"] = vec(H, 0.05, 1.0)
    if include_lm_head:
        w["lm_head.weight"] = mat(cfg.vocab, H, std=0.02)
    w["informative_head.weight"] = mat(2, H, std=0.02)
    w["relevance_head.weight"] = mat(2, H, std=0.02)
    return w


def synthetic_frames(n, seed=0, size=384, device="cuda"):
    """uint8 [n,3,size,size] smooth drifting sinusoid fields + 8-bit noise (a poor man's video), generated on `device`."""
    g = torch.Generator(device=device).manual_seed(seed)
    lin = torch.linspace(0, 1, size, device=device)
    yy, xx = torch.meshgrid(lin, lin, indexing="ij")
    ph = torch.rand(3, 3, generator=g, device=device) * 6.28
    fr = 1.0 + torch.rand(3, 3, generator=g, device=device) * 5.0
    t = torch.arange(n, device=device, dtype=torch.float32)[:, None, None, None]
    f = (torch.sin(6.28 * fr[:, 0, None, None] * xx + ph[:, 0, None, None] + 0.21 * t) *
         torch.cos(6.28 * fr[:, 1, None, None] * yy + ph[:, 1, None, None] - 0.13 * t) +
         0.5 * torch.sin(6.28 * fr[:, 2, None, None] * (xx + yy) + ph[:, 2, None, None] + 0.37 * t))
    img = f * 60.0 + 128.0 + torch.randn(n, 3, size, size, generator=g, device=device) * 6.0
    return img.clamp(0, 255).to(torch.uint8)
