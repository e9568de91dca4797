"""Tokenizer side of the path (models/tokenization_live.py:34-63,115-131): the chat template that frames system / user /
assistant / stream turns and the three prompt flags the loop uses (`add_stream_query_prompt`, `add_stream_prompt`,
`add_stream_generation_prompt`), plus the config update that records `v_placeholder_id` and `eos_token_id`.

`build_live_tokenizer_and_update_config(path, model_config)` takes a LOCAL tokenizer directory (there is no network for
`lmms-lab/llava-onevision-qwen2-7b-ov`); `SyntheticTokenizer` is a dependency-free stand-in with the same turn structure for
random-init benchmarks.  The loop only needs `apply_chat_template(..., return_tensors='pt')`, `decode`, `eos_token_id` and
`convert_tokens_to_ids`."""
import zlib

import torch

# Renders exactly what the reference's template renders (tests/test_tokenizer_cpu.py compares the two on every flag
# combination); written as one macro per turn instead of the reference's inline concatenations.
LIVE_CHAT_TEMPLATE = (
    "{%- macro turn(role, body) -%}{{ bos_token + role + '\\n' + body + eos_token }}{%- endmacro -%}"
    "{%- set ns = namespace(rest=messages) -%}"
    "{%- if messages[0]['role'] == 'system' -%}"
    "{{ turn('system', messages[0]['content']) }}{%- set ns.rest = messages[1:] -%}"
    "{%- endif -%}"
    "{%- for m in ns.rest -%}"
    "{%- if m['role'] == 'user' -%}{{ (eos_token if add_stream_query_prompt else '') + '\\n' + turn('user', m['content']) }}"
    "{%- elif m['role'] == 'assistant' -%}{{ '\\n' + turn('assistant', m['content']) }}"
    "{%- elif m['role'] == 'stream' and m['num_frames'] > 0 -%}{{ '\\n' + turn('stream', 'V_PLACEHOLDER' * (N_FRAME_TOKENS * m['num_frames'])) }}"
    "{%- endif -%}"
    "{%- endfor -%}"
    "{%- if add_generation_prompt -%}{{ '\\n' + bos_token + 'assistant\\n' }}"
    "{%- elif add_stream_prompt -%}{{ '\\n' + bos_token + 'stream\\n' }}"
    "{%- elif add_stream_generation_prompt -%}{{ eos_token + '\\n' + bos_token + 'assistant\\n' }}"
    "{%- endif -%}"
)


def live_chat_template(v_placeholder="<image>", frame_num_tokens=49):
    return LIVE_CHAT_TEMPLATE.replace("V_PLACEHOLDER", v_placeholder).replace("N_FRAME_TOKENS", str(int(frame_num_tokens)))


def build_live_tokenizer_and_update_config(llm_pretrained, model_config):
    """models/tokenization_live.py:115-131 for a local tokenizer directory: fast tokenizer, left padding, `<image>` registered
    as a special token, `<|im_start|>` / `<|im_end|>` as bos / eos, the live chat template; writes v_placeholder_id and
    eos_token_id into `model_config` (any object with attributes, or a dict)."""
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(llm_pretrained, use_fast=True, padding_side="left")
    get = (lambda k, d=None: model_config.get(k, d)) if isinstance(model_config, dict) else (lambda k, d=None: getattr(model_config, k, d))
    v = get("v_placeholder", "<image>") or "<image>"
    tok.add_special_tokens({"additional_special_tokens": [v]})
    tok.bos_token, tok.eos_token = "<|im_start|>", "<|im_end|>"
    tok.chat_template = live_chat_template(v, get("frame_num_tokens", 49) or 49)
    upd = dict(v_placeholder_id=tok.convert_tokens_to_ids(v), eos_token_id=tok.eos_token_id)
    if isinstance(model_config, dict):
        model_config.update(upd)
    else:
        for k, val in upd.items():
            setattr(model_config, k, val)
    return tok


class SyntheticTokenizer:
    """Deterministic token ids with the SAME turn structure the template produces (system turn; '\\n<|im_start|>stream\\n'
    stream prompt; user turn with the optional stream-query prefix; '<|im_end|>\\n<|im_start|>assistant\\n' generation prompt)
    for random-init benchmarks and fixtures: no tokenizer files needed."""

    IM_START, IM_END, NEWLINE, IMAGE = 3, 4, 5, 6      # IMAGE = the '<image>' frame placeholder (config.v_placeholder_id)

    def __init__(self, vocab_size, eos_token_id=None):
        self.vocab_size = vocab_size
        self.eos_token_id = self.IM_END if eos_token_id is None else eos_token_id

    def convert_tokens_to_ids(self, token):
        return {"<|im_start|>": self.IM_START, "<|im_end|>": self.IM_END, "<image>": self.IMAGE}.get(token)

    def _words(self, text):
        lo, span = 16, max(self.vocab_size - 16, 1)
        return [lo + zlib.crc32(w.encode()) % span for w in text.split()]

    def apply_chat_template(self, conversation, add_stream_prompt=False, add_stream_generation_prompt=False,
                            add_stream_query_prompt=False, return_tensors=None, **kw):
        ids = []
        for i, turn in enumerate(conversation):
            role = turn.get("role") if turn else None
            if role == "system":
                ids += [self.IM_START] + self._words("system") + [self.NEWLINE] + self._words(turn["content"]) + [self.IM_END]
            elif role == "user":
                if add_stream_query_prompt:
                    ids += [self.IM_END, self.NEWLINE]
                ids += [self.IM_START] + self._words("user") + [self.NEWLINE] + self._words(turn["content"]) + [self.IM_END]
            elif role == "assistant":
                ids += [self.IM_START] + self._words("assistant") + [self.NEWLINE] + self._words(turn["content"]) + [self.IM_END]
        if add_stream_prompt:
            ids += [self.NEWLINE, self.IM_START] + self._words("stream") + [self.NEWLINE]
        if add_stream_generation_prompt:
            ids += [self.IM_END, self.NEWLINE, self.IM_START] + self._words("assistant") + [self.NEWLINE]
        t = torch.tensor([ids], dtype=torch.long)
        return t if return_tensors == "pt" else ids

    def decode(self, ids, skip_special_tokens=True, clean_up_tokenization_spaces=True):
        ids = ids.tolist() if torch.is_tensor(ids) else list(ids)
        if skip_special_tokens:
            ids = [i for i in ids if i not in (self.IM_START, self.IM_END, self.NEWLINE)]
        return " ".join(f"<{i}>" for i in ids)
