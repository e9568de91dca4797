"""Stand-in for the reference's chat-template tokenizer (models/tokenization_live.py:34-63,115-131).

The Qwen2 tokenizer files cannot be fetched offline, so `SyntheticTokenizer` produces deterministic token ids with the
SAME structure the reference's Jinja template produces (system turn; '\\n<|im_start|>stream\\n' stream prompt; user turn
with optional stream-query prefix; '<|im_end|>\\n<|im_start|>assistant\\n' generation prompt).  A real HF tokenizer with
the reference's template can be passed to LiveInferForBenchmark instead; only `apply_chat_template(...,
return_tensors='pt')` and `decode` are used."""
import zlib

import torch


class SyntheticTokenizer:
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
