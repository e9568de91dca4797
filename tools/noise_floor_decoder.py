"""Where does the decoder's score error come from?  The fp32 oracle vs torch emulations of the same decoder that round to
bf16 at ONE of the kernel design's rounding sites at a time (and at all of them), vs the CUDA decoder, on the decoder pass
shapes the parity tests use.  Writes gpurun_out/noise_floor_decoder.json.  (Test-side tool: imports oracle/.)"""
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
SITES = ["x_attn", "q", "k", "v", "p", "attn_out", "x_mlp", "h"]


def rb(x, on):
    return x.bfloat16().float() if on else x


def decoder_emul(w, arch, x, sites=(), hilo=(), exact_rows=None, exact_sites=()):
    """R.decoder_forward for one stream from an empty cache, rounding the named intermediate tensors to bf16.
    `hilo`: sites whose rounding is replaced by a hi+lo bf16 pair (error 2^-17 instead of 2^-9)."""
    def r(t, name):
        if exact_rows is not None and name in exact_sites and name in sites:
            # rounded everywhere except on `exact_rows` (hi+lo there): the "precise score rows" design
            tt = t.transpose(0, 1) if name in ("q", "k", "v") else t
            out = tt.bfloat16().float()
            hi = tt[exact_rows].bfloat16().float()
            out[exact_rows] = hi + (tt[exact_rows] - hi).bfloat16().float()
            return out.transpose(0, 1) if name in ("q", "k", "v") else out
        if name in hilo:
            hi = t.bfloat16().float()
            return hi + (t - hi).bfloat16().float()
        return rb(t, name in sites)
    M = x.shape[0]
    pos = torch.arange(M, device=x.device)
    cos, sin = R.rope_cos_sin(arch, pos, x.dtype)
    Hq, Hkv, dh = arch.q_heads, arch.kv_heads, arch.head_dim
    mask = torch.ones(M, M, dtype=torch.bool, device=x.device).tril()
    for i in range(arch.layers):
        p = f"model.layers.{i}."
        h = r(R.rms_norm(x, w[p + "input_layernorm.weight"], arch.rms_eps), "x_attn")
        q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]).view(M, Hq, dh).transpose(0, 1)
        k = F.linear(h, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"]).view(M, Hkv, dh).transpose(0, 1)
        v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]).view(M, Hkv, dh).transpose(0, 1)
        q = r(q * cos[None] + R.rotate_half(q) * sin[None], "q")
        k = r(k * cos[None] + R.rotate_half(k) * sin[None], "k")
        v = r(v, "v")
        kk = k.repeat_interleave(Hq // Hkv, dim=0)
        vv = v.repeat_interleave(Hq // Hkv, dim=0)
        s = (q @ kk.transpose(-1, -2)) * dh ** -0.5
        s = s.masked_fill(~mask[None], float("-inf"))
        a = r(torch.softmax(s, dim=-1), "p")
        del s
        o = r((a @ vv).transpose(0, 1).reshape(M, Hq * dh), "attn_out")
        del a
        x = x + F.linear(o, w[p + "self_attn.o_proj.weight"])
        h = r(R.rms_norm(x, w[p + "post_attention_layernorm.weight"], arch.rms_eps), "x_mlp")
        h = r(F.silu(F.linear(h, w[p + "mlp.gate_proj.weight"])) * F.linear(h, w[p + "mlp.up_proj.weight"]), "h")
        x = x + F.linear(h, w[p + "mlp.down_proj.weight"])
    return R.rms_norm(x, w["model.norm.weight"], arch.rms_eps)


def scores_of(w, hidden, rows):
    il = F.linear(hidden[rows], w["informative_head.weight"])
    rl = F.linear(hidden[rows], w["relevance_head.weight"])
    return torch.stack([il.softmax(-1)[:, 1], rl.softmax(-1)[:, 1]], 1), torch.cat([il, rl], 1)


def main():
    from mmduet_b200.config import ModelConfig
    from mmduet_b200.engine import DecoderEngine
    arch = A.FULL
    w = R.make_weights(arch, seed=1234, device=dev, generate_on_device=True, include_lm_head=False)
    w = {k: v for k, v in w.items() if "vision_tower" not in k}
    torch.cuda.empty_cache()
    dec = DecoderEngine(ModelConfig.from_any(arch), w, dev, max_context=4096, max_tokens=2048)
    rep = {}
    for seed, k in ((9, 24), (109, 24), (140, 40)):
        g = torch.Generator(device=dev).manual_seed(seed)
        fe = (torch.randn(k * 49, arch.hidden, generator=g, device=dev) * 1.14).bfloat16()
        rows = torch.tensor([49 * (j + 1) - 1 for j in range(k)], device=dev)
        x = fe.float()
        ref_h = decoder_emul(w, arch, x)
        ref_s, ref_l = scores_of(w, ref_h, rows)
        case = {"hidden_rms": ref_h.pow(2).mean().sqrt().item(), "logit_std": ref_l.std().item()}

        def err(sites=(), hilo=(), exact_rows=None, exact_sites=()):
            h = decoder_emul(w, arch, x, sites, hilo, exact_rows, exact_sites)
            s, l = scores_of(w, h, rows)
            return {"score_maxabs": (s - ref_s).abs().max().item(), "score_rms": (s - ref_s).pow(2).mean().sqrt().item(),
                    "logit_maxabs": (l - ref_l).abs().max().item(), "hidden_maxabs": (h[rows] - ref_h[rows]).abs().max().item()}
        if os.environ.get("NFD_ALL"):
            for sname in SITES:
                case["only_" + sname] = err((sname,))
        case["all_sites"] = err(tuple(SITES))
        for combo in (("h", "x_mlp"), ("h", "x_mlp", "x_attn")):
            case["all_but_hilo_" + "+".join(combo)] = err(tuple(SITES), combo)
        for es in (("x_mlp",), ("x_mlp", "h"), ("x_mlp", "h", "x_attn"), ("x_mlp", "h", "x_attn", "attn_out")):
            case["score_rows_exact_" + "+".join(es)] = err(tuple(SITES), (), rows, es)
        st = dec.new_stream()
        out = dec.step([dict(storage=st, past=0, ids=[], frames=fe, score_rows=rows.tolist())], score="frame_ends")
        case["cuda_pass"] = {"score_maxabs": (out["scores"] - ref_s).abs().max().item(),
                             "score_rms": (out["scores"] - ref_s).pow(2).mean().sqrt().item(),
                             "logit_maxabs": (out["head_logits"] - ref_l).abs().max().item()}
        st.release()
        rep[f"seed{seed}_k{k}"] = case
        print(f"seed{seed}_k{k}", json.dumps(case, indent=1), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rep, open("gpurun_out/noise_floor_decoder.json", "w"), indent=1)


if __name__ == "__main__":
    main()
