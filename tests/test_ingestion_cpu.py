"""CPU tests of weight / tokenizer ingestion: safetensors checkpoints and peft adapter directories,
LoRA merge incl. modules_to_save, and a real `PreTrainedTokenizerFast` built offline that carries the live chat template."""
import json
import os

import pytest
import torch


def offline_tokenizer_dir(path):
    """A genuine fast tokenizer without any download: byte-level BPE with an empty merge table (256 byte symbols), the two
    ChatML specials of Qwen2.  Saved with save_pretrained so that AutoTokenizer.from_pretrained(path) loads it."""
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    alphabet = sorted(pre_tokenizers.ByteLevel.alphabet())
    tok = Tokenizer(models.BPE(vocab={c: i for i, c in enumerate(alphabet)}, merges=[]))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<|im_start|>", eos_token="<|im_end|>")
    fast.save_pretrained(path)
    return path


CONVERSATIONS = [
    ([{"role": "system", "content": "You watch a video."}], {}),
    ([{}], dict(add_stream_prompt=True)),
    ([{}], dict(add_stream_generation_prompt=True)),
    ([{"role": "user", "content": "what now?"}], dict(add_stream_query_prompt=True, add_stream_prompt=True)),
    ([{"role": "user", "content": "what now?"}], dict(add_stream_query_prompt=False, add_stream_prompt=True)),
    ([{"role": "system", "content": "S"}, {"role": "stream", "num_frames": 2}, {"role": "user", "content": "Q"},
      {"role": "assistant", "content": "A"}, {"role": "stream", "num_frames": 0}, {"role": "stream", "num_frames": 1}], dict(add_generation_prompt=True)),
]


def test_live_tokenizer_from_local_directory(tmp_path):
    from mmduet_b200.inference import template_ids
    from mmduet_b200.tokenization_live import build_live_tokenizer_and_update_config
    cfg = dict(v_placeholder="<image>", frame_num_tokens=3)
    tok = build_live_tokenizer_and_update_config(offline_tokenizer_dir(str(tmp_path / "tok")), cfg)
    assert cfg["eos_token_id"] == tok.eos_token_id == tok.convert_tokens_to_ids("<|im_end|>")
    assert cfg["v_placeholder_id"] == tok.convert_tokens_to_ids("<image>") and cfg["v_placeholder_id"] >= 256
    text = tok.apply_chat_template(CONVERSATIONS[5][0], tokenize=False, **CONVERSATIONS[5][1])
    assert text == ("<|im_start|>system\nS<|im_end|>\n<|im_start|>stream\n" + "<image>" * 6 + "<|im_end|>\n<|im_start|>user\nQ<|im_end|>"
                    "\n<|im_start|>assistant\nA<|im_end|>\n<|im_start|>stream\n<image><image><image><|im_end|>\n<|im_start|>assistant\n")
    assert tok.apply_chat_template([{}], tokenize=False, add_stream_prompt=True) == "\n<|im_start|>stream\n"
    assert tok.apply_chat_template([{}], tokenize=False, add_stream_generation_prompt=True) == "<|im_end|>\n<|im_start|>assistant\n"
    q = tok.apply_chat_template([{"role": "user", "content": "hi"}], tokenize=False, add_stream_query_prompt=True, add_stream_prompt=True)
    assert q == "<|im_end|>\n<|im_start|>user\nhi<|im_end|>\n<|im_start|>stream\n"
    ids = template_ids(tok, [{"role": "user", "content": "hi"}], add_stream_query_prompt=True, add_stream_prompt=True)
    assert ids[0] == tok.eos_token_id and ids.count(tok.bos_token_id) == 2 and tok.decode(ids, skip_special_tokens=True) == "\nuser\nhi\nstream\n"
    assert ids.count(cfg["v_placeholder_id"]) == 0
    frame_ids = template_ids(tok, [{"role": "stream", "num_frames": 2}])
    assert frame_ids.count(cfg["v_placeholder_id"]) == 6           # '<image>' is one token each


@pytest.mark.reference
def test_live_template_renders_like_the_reference_template(tmp_path):
    """Our template against the reference's own template string (models/tokenization_live.py:34-63) on the same tokenizer."""
    from mmduet_b200.tokenization_live import build_live_tokenizer_and_update_config
    from oracle import ref_import as RI
    RI.import_reference()
    import importlib
    tl = importlib.import_module("models.tokenization_live")

    class Cfg:
        frame_num_tokens, v_placeholder = 5, "<image>"
    ours = build_live_tokenizer_and_update_config(offline_tokenizer_dir(str(tmp_path / "a")), dict(frame_num_tokens=5))
    theirs = build_live_tokenizer_and_update_config(offline_tokenizer_dir(str(tmp_path / "b")), dict(frame_num_tokens=5))
    theirs.chat_template = tl.chat_template_llava(theirs, tl.get_stream_placeholder_jinja2(Cfg))
    for conv, flags in CONVERSATIONS:
        assert ours.apply_chat_template(conv, tokenize=False, **flags) == theirs.apply_chat_template(conv, tokenize=False, **flags), (conv, flags)


def test_safetensors_checkpoint_and_adapter_directories(tmp_path):
    from safetensors.torch import save_file
    from mmduet_b200.checkpoint import load_safetensors, merge_lora, state_dict_from_pretrained
    torch.manual_seed(0)
    W, A, B = torch.randn(8, 6).bfloat16(), torch.randn(4, 6).bfloat16(), torch.randn(8, 4).bfloat16()
    head, head_ft = torch.randn(2, 6).bfloat16(), torch.randn(2, 6).bfloat16()
    base = {"model.layers.0.self_attn.q_proj.weight": W, "informative_head.weight": head, "model.norm.weight": torch.ones(6)}
    ck = tmp_path / "ckpt"
    ck.mkdir()
    save_file({k: v for k, v in list(base.items())[:2]}, str(ck / "model-00001-of-00002.safetensors"))
    save_file({"model.norm.weight": base["model.norm.weight"]}, str(ck / "model-00002-of-00002.safetensors"))
    json.dump({"weight_map": {"model.layers.0.self_attn.q_proj.weight": "model-00001-of-00002.safetensors",
                              "informative_head.weight": "model-00001-of-00002.safetensors",
                              "model.norm.weight": "model-00002-of-00002.safetensors"}}, open(ck / "model.safetensors.index.json", "w"))
    sd, _ = load_safetensors(str(ck))
    assert set(sd) == set(base) and all(torch.equal(sd[k], base[k]) for k in base)
    ad = tmp_path / "adapter"
    ad.mkdir()
    save_file({"base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight": A,
               "base_model.model.model.layers.0.self_attn.q_proj.lora_B.weight": B,
               "base_model.model.informative_head.weight": head_ft}, str(ad / "adapter_model.safetensors"))    # modules_to_save copy
    json.dump({"r": 4, "lora_alpha": 8}, open(ad / "adapter_config.json", "w"))
    merged = state_dict_from_pretrained(str(ck), str(ad))
    want = (W.float() + 2.0 * (B.float() @ A.float())).bfloat16()
    assert torch.equal(merged["model.layers.0.self_attn.q_proj.weight"], want)
    assert torch.equal(merged["informative_head.weight"], head_ft) and torch.equal(merged["model.norm.weight"], base["model.norm.weight"])
    # a live PeftModel.state_dict(): base_layer / lora_X.default / modules_to_save.default / original_module
    live = {"base_model.model.model.layers.0.self_attn.q_proj.base_layer.weight": W,
            "base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight": A,
            "base_model.model.model.layers.0.self_attn.q_proj.lora_B.default.weight": B,
            "base_model.model.informative_head.original_module.weight": head,
            "base_model.model.informative_head.modules_to_save.default.weight": head_ft,
            "base_model.model.model.norm.weight": base["model.norm.weight"]}
    m2 = merge_lora(live, None, lora_r=4, lora_alpha=8)
    assert set(m2) == set(base) and torch.equal(m2["informative_head.weight"], head_ft)
    assert torch.equal(m2["model.layers.0.self_attn.q_proj.weight"], want)
    with pytest.raises(FileNotFoundError):
        load_safetensors(str(tmp_path / "lmms-lab" / "llava-onevision-qwen2-7b-ov"))
