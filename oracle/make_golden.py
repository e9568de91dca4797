"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the reference's OWN classes
(VideoHeadLiveLlavaQwenForCausalLM.visual_embed / .forward and models/vision_live._siglip_vision_encode, imported
from /root/reference via oracle/ref_import.py) on seeded inputs.  Run in the authoring container:

    python -m oracle.make_golden

The fixtures pin oracle/restate.py (CPU tests) and the CUDA path (GPU tests) to the reference without needing
/root/reference at test time.  Weights are not stored: they are regenerated from the seed by restate.make_weights.
"""
import os

import numpy as np
import torch

from . import arch as A
from . import ref_import as RI
from . import restate as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_stream(arch, name, seed, n_frames, prefix_len, pool_mode="bilinear"):
    from transformers import DynamicCache
    arch = A.with_pool(arch, pool_mode)
    w = R.make_weights(arch, seed=seed)
    model = RI.build_reference_model(arch, state_dict=w, dtype=torch.float32)
    frames = R.synthetic_frames(n_frames, seed=seed + 1)
    px = R.preprocess_frames(frames)
    emb = model.visual_embed(px)                                  # models/modeling_live.py:26-33
    g = torch.Generator().manual_seed(seed + 2)
    prefix = torch.randint(0, arch.vocab, (prefix_len,), generator=g)
    cache = DynamicCache(config=model.config)
    inf, rel, last_logits = [], [], []
    tpf = emb.shape[0] // n_frames      # 49 for bilinear; avg/max pool(27, stride 4) floors to 6x6 = 36
    for f in range(n_frames):
        fe = emb[f * tpf:(f + 1) * tpf]
        pre = model.get_input_embeddings()(prefix) if f == 0 else torch.zeros(0, arch.hidden)
        x = torch.cat([pre, fe])[None]
        out = model(inputs_embeds=x, use_cache=True, past_key_values=cache, return_dict=True)   # test/inference.py:239
        cache = out.past_key_values
        inf.append(out.informative_logits[0, -1].numpy())
        rel.append(out.relevance_logits[0, -1].numpy())
        last_logits.append(out.logits[0, -1].numpy())
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), seed=seed, n_frames=n_frames, prefix=prefix.numpy(),
        frame_embeds=emb.numpy().astype(np.float32), informative_logits=np.stack(inf), relevance_logits=np.stack(rel),
        lm_logits_last=np.stack(last_logits).astype(np.float32), pool_mode=pool_mode, tokens_per_frame=tpf)
    print(name, emb.shape, np.stack(inf)[:, 1] - np.stack(inf)[:, 0])


def golden_legacy(arch, name, seed, n_frames):
    """models/vision_live.py:11-31 run unmodified on an HF SiglipVisionModel.vision_model."""
    from transformers import SiglipVisionConfig, SiglipVisionModel
    vl, _, _ = RI.import_reference()
    w = R.make_weights(arch, seed=seed, legacy_post_ln=True)
    vcfg = SiglipVisionConfig(hidden_size=arch.vit_dim, intermediate_size=arch.vit_mlp,
                              num_hidden_layers=arch.vit_layers_total, num_attention_heads=arch.vit_heads,
                              image_size=arch.image_size, patch_size=arch.patch_size, hidden_act="gelu_pytorch_tanh",
                              layer_norm_eps=1e-6)
    vm = SiglipVisionModel(vcfg).vision_model.eval()
    sd = {k[len(R.VT):]: v for k, v in w.items() if k.startswith(R.VT)}
    missing, unexpected = vm.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    frames = R.synthetic_frames(n_frames, seed=seed + 1).float()
    with torch.no_grad():
        out = vl._siglip_vision_encode(vm, frames, frame_token_cls=False, frame_token_pooled=(7, 7))
        cls_sp = vl._siglip_vision_encode(vm, frames, frame_token_cls=True, frame_token_pooled=(3, 3))     # [T, 1 + 9, D]
        cls_only = vl._siglip_vision_encode(vm, frames, frame_token_cls=True, frame_token_pooled=None)      # [T, 1, D]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, n_frames=n_frames, tokens=out.numpy(), cls_tokens=cls_sp.numpy(),
                        cls_only=cls_only.numpy())
    print(name, out.shape, cls_sp.shape, cls_only.shape)


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    torch.set_num_threads(8)
    if "--legacy-only" not in sys.argv:
        golden_stream(A.TINY, "tiny_stream_bilinear", seed=11, n_frames=4, prefix_len=9)
        golden_stream(A.TINY, "tiny_stream_average", seed=12, n_frames=2, prefix_len=5, pool_mode="average")
        golden_stream(A.TINY, "tiny_stream_max", seed=13, n_frames=2, prefix_len=5, pool_mode="max")
        golden_stream(A.SMALL, "small_stream_bilinear", seed=21, n_frames=3, prefix_len=13)
    golden_legacy(A.TINY, "tiny_legacy_vision", seed=31, n_frames=2)


if __name__ == "__main__":
    main()
