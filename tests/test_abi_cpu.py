"""CPU tests of the boundary: the shared library loads, exports every symbol include/mmduet_b200.h declares, and the host
logic that needs no GPU (tokenizer stand-in, argument dataclass, pooling tap tables) behaves like the reference."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mmduet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mmduet_b200.h")).read()
    declared = set(re.findall(r"MMD_API\s+[\w\s\*]+?\b(mmd_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert b"sm_100a" in lib.mmd_version()


def test_no_cpu_fallback():
    from mmduet_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert not lib.mmd_create(0)                      # no device: NULL, with a message
    assert lib.mmd_last_error() != b""
    import mmduet_b200
    with pytest.raises(ValueError):
        mmduet_b200.build_model_and_tokenizer(is_training=False)   # no state_dict, nothing silently substituted


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mmduet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_pooling_taps_match_reference_ops():
    import torch.nn.functional as F
    from mmduet_b200.engine import pooling_taps, taps_to_tables
    x = torch.randn(2, 729, 5)
    xi = x.view(2, 27, 27, 5).permute(0, 3, 1, 2)
    for mode, ref in (("bilinear", F.interpolate(xi, size=[7, 7], mode="bilinear")), ("average", F.avg_pool2d(xi, 4))):
        taps = pooling_taps(27, 4, mode)
        got = torch.einsum("os,tsd->tod", taps, x)
        assert (got - ref.permute(0, 2, 3, 1).reshape(2, -1, 5)).abs().max() < 1e-5
        gidx, tidx, tw, mt = taps_to_tables(taps)
        assert gidx.numel() == (169 if mode == "bilinear" else 576) and tidx.shape == (taps.shape[0], mt)
    assert pooling_taps(27, 4, "bilinear").shape[0] == 49 and pooling_taps(27, 4, "max").shape[0] == 36
    with pytest.raises(ValueError):
        pooling_taps(27, 4, "nearest")


def test_tokenizer_and_arguments_shapes():
    from mmduet_b200.arguments_live import LiveTestArguments
    from mmduet_b200.tokenization_live import SyntheticTokenizer
    a = LiveTestArguments()
    assert a.frame_num_tokens == 49 and a.video_pooling_stride == 4 and a.score_heads == "informative_score"
    assert a.running_list_length == 20 and a.remove_assistant_turns is False and a.frame_fps == 2
    t = SyntheticTokenizer(152064)
    start = t.apply_chat_template([{"role": "system", "content": a.system_prompt}], return_tensors="pt")
    assert start.shape[0] == 1 and start.dtype == torch.long and 20 < start.shape[1] < 64
    plain = t.apply_chat_template([{"role": "user", "content": "hi there"}], add_stream_prompt=True, return_tensors="pt")
    stream_q = t.apply_chat_template([{"role": "user", "content": "hi there"}], add_stream_query_prompt=True, add_stream_prompt=True, return_tensors="pt")
    assert stream_q.shape[1] == plain.shape[1] + 2          # '<|im_end|>\\n' closes the stream turn first
    assert int(start.max()) < 152064


def test_config_validation():
    from mmduet_b200.config import ModelConfig
    ModelConfig().validate()
    with pytest.raises(ValueError):
        ModelConfig(vit_dim=1024).validate()                 # head_dim 64: no kernel instantiation
    with pytest.raises(ValueError):
        ModelConfig(pool_mode="nearest").validate()


def test_merge_lora_matches_peft_formula():
    from mmduet_b200.checkpoint import merge_lora
    g = torch.Generator().manual_seed(0)
    W = torch.randn(12, 8, generator=g).bfloat16()
    A = torch.randn(4, 8, generator=g) * 0.1
    B = torch.randn(12, 4, generator=g) * 0.1
    base = {"model.layers.0.self_attn.q_proj.weight": W, "model.norm.weight": torch.ones(8)}
    lora = {"base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight": A,
            "base_model.model.model.layers.0.self_attn.q_proj.lora_B.weight": B}
    merged = merge_lora(base, lora, lora_r=4, lora_alpha=8)
    want = (W.float() + 2.0 * (B @ A)).bfloat16()
    assert torch.equal(merged["model.layers.0.self_attn.q_proj.weight"], want)
    assert torch.equal(merged["model.norm.weight"], base["model.norm.weight"])
    # single dict in peft's in-model layout (base_layer + lora_A/B.default)
    combo = {"base_model.model.model.layers.0.self_attn.q_proj.base_layer.weight": W,
             "base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight": A,
             "base_model.model.model.layers.0.self_attn.q_proj.lora_B.default.weight": B}
    merged2 = merge_lora(combo, None, lora_r=4, lora_alpha=8)
    assert torch.equal(merged2["model.layers.0.self_attn.q_proj.weight"], want)
    with pytest.raises(KeyError):
        merge_lora({"x.weight": W}, lora, 4, 8)
