"""Where does the ViT error come from?  fp32 oracle vs torch emulations that round at the kernel design's rounding
sites, vs the CUDA tower.  Writes gpurun_out/noise_floor.json."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import arch as A  # noqa: E402
from oracle import restate as R  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
VT = R.VT


def rb(x, on):
    return x.bfloat16().float() if on else x


def tower_emul(w, arch, px, ln=False, qkv=False, p=False, attn=False, h=False, pix=False, n_layers=None):
    x = R.siglip_embeddings(w, arch, rb(px, pix))
    outs = []
    for i in range(arch.vit_layers if n_layers is None else n_layers):
        pf = f"{VT}encoder.layers.{i}."
        T, S, D = x.shape
        H, dh = arch.vit_heads, arch.vit_head_dim
        hh = rb(F.layer_norm(x, (D,), w[pf + "layer_norm1.weight"], w[pf + "layer_norm1.bias"], 1e-6), ln)
        q = rb(F.linear(hh, w[pf + "self_attn.q_proj.weight"], w[pf + "self_attn.q_proj.bias"]), qkv).view(T, S, H, dh).transpose(1, 2)
        k = rb(F.linear(hh, w[pf + "self_attn.k_proj.weight"], w[pf + "self_attn.k_proj.bias"]), qkv).view(T, S, H, dh).transpose(1, 2)
        v = rb(F.linear(hh, w[pf + "self_attn.v_proj.weight"], w[pf + "self_attn.v_proj.bias"]), qkv).view(T, S, H, dh).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) * dh ** -0.5, dim=-1)
        o = rb((rb(att, p) @ v).transpose(1, 2).reshape(T, S, D), attn)
        x = x + F.linear(o, w[pf + "self_attn.out_proj.weight"], w[pf + "self_attn.out_proj.bias"])
        hh = rb(F.layer_norm(x, (D,), w[pf + "layer_norm2.weight"], w[pf + "layer_norm2.bias"], 1e-6), ln)
        hh = rb(F.gelu(F.linear(hh, w[pf + "mlp.fc1.weight"], w[pf + "mlp.fc1.bias"]), approximate="tanh"), h)
        x = x + F.linear(hh, w[pf + "mlp.fc2.weight"], w[pf + "mlp.fc2.bias"])
        outs.append(x)
    return x, outs


def main():
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.engine import VisionEngine
    arch = A.FULL
    w = R.make_weights(arch, seed=1234, device=dev, generate_on_device=True, include_lm_head=False)
    w = {k: v for k, v in w.items() if not k.startswith("model.layers")}
    torch.cuda.empty_cache()
    px = R.preprocess_frames(R.synthetic_frames(2, seed=8)).bfloat16().float().to(dev)
    rep = {}
    ref, ref_layers = tower_emul(w, arch, px)
    rep["ref_absmax"] = ref.abs().max().item()
    rep["ref_rms"] = ref.pow(2).mean().sqrt().item()

    def err(a, b):
        d = (a - b).abs()
        return {"max": d.max().item(), "rms": d.pow(2).mean().sqrt().item()}
    for name, kw in [("pix", dict(pix=True)), ("ln", dict(ln=True)), ("qkv", dict(qkv=True)), ("p", dict(p=True)), ("attn", dict(attn=True)),
                     ("h", dict(h=True)), ("all", dict(ln=True, qkv=True, p=True, attn=True, h=True, pix=True))]:
        out, layers = tower_emul(w, arch, px, **kw)
        rep["emul_" + name] = err(out, ref)
        if name == "all":
            emul_all, emul_layers = out, layers
            rep["emul_all_by_layer_rms"] = [err(a, b)["rms"] for a, b in zip(layers, ref_layers)]
    vis = VisionEngine(ModelConfig.from_any(arch), w, dev)
    hid = vis.tower(px).view(2, arch.patches, arch.vit_dim).clone()
    rep["cuda_vs_ref"] = err(hid, ref)
    rep["cuda_vs_emul_all"] = err(hid, emul_all)
    by_layer = []
    for nl in (1, 2, 4, 8, 16, 26):
        v2 = VisionEngine(ModelConfig.from_any(arch), w, dev, with_projector=False, n_layers=nl)
        hl = v2.tower(px).view(2, arch.patches, arch.vit_dim)
        by_layer.append({"layers": nl, "cuda_vs_ref": err(hl, ref_layers[nl - 1]), "cuda_vs_emul": err(hl, emul_layers[nl - 1]),
                         "emul_vs_ref": err(emul_layers[nl - 1], ref_layers[nl - 1])})
        del v2
    rep["by_layer"] = by_layer
    # embeddings
    def proj(hh, rnd):
        g = rb(hh, rnd)
        a = rb(F.gelu(F.linear(g, w["model.mm_projector.0.weight"], w["model.mm_projector.0.bias"])), rnd)
        y = F.linear(a, w["model.mm_projector.2.weight"], w["model.mm_projector.2.bias"])
        return R.post_projector_pooling(arch, y).view(-1, arch.hidden)
    e_ref = proj(ref, False)
    rep["emb_ref_absmax"] = e_ref.abs().max().item()
    rep["emb_emul_all_prebf16"] = err(proj(emul_all, True), e_ref)
    rep["emb_emul_all_bf16"] = err(proj(emul_all, True).bfloat16().float(), e_ref)
    rep["emb_ref_hidden_proj_rounded_bf16"] = err(proj(ref, True).bfloat16().float(), e_ref)
    rep["emb_ref_bf16_only"] = err(e_ref.bfloat16().float(), e_ref)
    emb = vis.visual_embed(px)
    rep["emb_cuda"] = err(emb.float(), e_ref)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rep, open("gpurun_out/noise_floor.json", "w"), indent=1)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
