"""TEST INFRASTRUCTURE ONLY — CPU/torch restatement of MMDuet's per-frame streaming hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import oracle/; the
product (mmduet_b200/) never does.  PARITY PINNING: the reference ships no golden vectors or known-answer tests for
this path, so this restatement is pinned against outputs of the reference's OWN classes imported
from /root/reference through oracle/ref_import.py (tests/test_oracle.py::test_restatement_matches_reference_import in
the authoring container) and against the fixtures those classes generated (tests/golden/, oracle/make_golden.py).

Every function is plain functional torch over a flat weight dict that uses the reference's state_dict key names
(`model.vision_tower.vision_tower.vision_model...`, `model.mm_projector.{0,2}`, `model.layers.N...`, `model.norm`,
`lm_head`, `informative_head`, `relevance_head`).  "Oracle of record" = this code run in fp32 on bf16-rounded
weights and inputs (SURVEY.md §8c tolerances).

Third-party arithmetic restated here (absent from /root/reference): LLaVA-NeXT `llava` (git HEAD, unpinned:
SigLipVisionTower, mlp2x_gelu projector) and transformers==4.44.2 (Qwen2, SigLIP); the restatement follows the
installed transformers 5.5.0 sources (TF:) and the 4.44.2 cache-rollback meaning described in SURVEY.md §3.3.
"""
import math

import torch
import torch.nn.functional as F

VT = "model.vision_tower.vision_tower.vision_model."


# ------------------------------------------------------------------------------------------------------------
# SigLIP tower  (TF:models/siglip/modeling_siglip.py:116-186 embeddings, :252-362 attention/MLP/layer)
# ------------------------------------------------------------------------------------------------------------
def siglip_embeddings(w, arch, pixels):
    """Conv2d(3, D, k=14, s=14, 'valid') + learned position embedding -> [T, 729, D]."""
    x = F.conv2d(pixels, w[VT + "embeddings.patch_embedding.weight"], w[VT + "embeddings.patch_embedding.bias"],
                 stride=arch.patch_size)
    x = x.flatten(2).transpose(1, 2)
    return x + w[VT + "embeddings.position_embedding.weight"][None]


def siglip_layer(w, arch, i, x):
    p = f"{VT}encoder.layers.{i}."
    T, S, D = x.shape
    H, dh = arch.vit_heads, arch.vit_head_dim
    h = F.layer_norm(x, (D,), w[p + "layer_norm1.weight"], w[p + "layer_norm1.bias"], 1e-6)
    q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]).view(T, S, H, dh).transpose(1, 2)
    k = F.linear(h, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"]).view(T, S, H, dh).transpose(1, 2)
    v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]).view(T, S, H, dh).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * dh ** -0.5, dim=-1, dtype=torch.float32).to(q.dtype)
    o = (att @ v).transpose(1, 2).reshape(T, S, D)
    x = x + F.linear(o, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
    h = F.layer_norm(x, (D,), w[p + "layer_norm2.weight"], w[p + "layer_norm2.bias"], 1e-6)
    h = F.gelu(F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"]), approximate="tanh")
    return x + F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])


def siglip_tower(w, arch, pixels, n_layers=None):
    """llava path (video_head_live_llava_qwen.py:96-98 -> LLaVA SigLipVisionTower): `vit_layers` layers, NO post-LN."""
    x = siglip_embeddings(w, arch, pixels)
    for i in range(arch.vit_layers if n_layers is None else n_layers):
        x = siglip_layer(w, arch, i, x)
    return x


# ------------------------------------------------------------------------------------------------------------
# projector + pooling  (video_head_live_llava_qwen.py:90-91, :100-119)
# ------------------------------------------------------------------------------------------------------------
def connector(w, feats):
    h = F.gelu(F.linear(feats, w["model.mm_projector.0.weight"], w["model.mm_projector.0.bias"]))
    return F.linear(h, w["model.mm_projector.2.weight"], w["model.mm_projector.2.bias"])


def post_projector_pooling(arch, x):
    T, S, D = x.shape
    g = arch.grid
    x = x.view(T, g, g, D).permute(0, 3, 1, 2).contiguous()
    if arch.pool_mode == "average":
        x = F.avg_pool2d(x, arch.pool_stride)
    elif arch.pool_mode == "max":
        x = F.max_pool2d(x, arch.pool_stride)
    elif arch.pool_mode == "bilinear":
        size = [math.ceil(g / arch.pool_stride)] * 2
        x = F.interpolate(x, size=size, mode="bilinear")
    else:
        raise ValueError(f"Unexpected mm_spatial_pool_mode: {arch.pool_mode}")
    return x.permute(0, 2, 3, 1).reshape(T, -1, D).contiguous()


def visual_embed(w, arch, pixels):
    """LiveMixin.visual_embed (models/modeling_live.py:26-33): tower -> connector -> pooling -> [T*tokens, hidden]."""
    feats = siglip_tower(w, arch, pixels)
    x = post_projector_pooling(arch, connector(w, feats))
    return x.view(-1, x.shape[-1])


def preprocess_frames(frames_u8):
    """LLaVA SigLipImageProcessor on already-384x384 frames: x/255 then (x-0.5)/0.5 (resize is the identity)."""
    return (frames_u8.float() * 0.00392156862745098 - 0.5) / 0.5


# ------------------------------------------------------------------------------------------------------------
# legacy entry: models/vision_live.py:11-31 (_siglip_vision_encode on an HF SiglipVisionModel.vision_model)
# ------------------------------------------------------------------------------------------------------------
def siglip_pooling_head(w, arch, h):
    """SiglipMultiheadAttentionPoolingHead (TF:models/siglip/modeling_siglip.py): a learned probe attends over the patch tokens
    (nn.MultiheadAttention, batch_first), then x + MLP(LayerNorm(x)); returns [T, D] = vision_outputs.pooler_output."""
    hp = VT + "head."
    T, S, D = h.shape
    H, dh = arch.vit_heads, arch.vit_head_dim
    w_in, b_in = w[hp + "attention.in_proj_weight"], w[hp + "attention.in_proj_bias"]
    q = F.linear(w[hp + "probe"].expand(T, 1, D), w_in[:D], b_in[:D]).view(T, 1, H, dh).transpose(1, 2)
    k = F.linear(h, w_in[D:2 * D], b_in[D:2 * D]).view(T, S, H, dh).transpose(1, 2)
    v = F.linear(h, w_in[2 * D:], b_in[2 * D:]).view(T, S, H, dh).transpose(1, 2)
    a = torch.softmax((q * dh ** -0.5) @ k.transpose(-1, -2), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(T, 1, D)
    x = F.linear(o, w[hp + "attention.out_proj.weight"], w[hp + "attention.out_proj.bias"])
    y = F.layer_norm(x, (D,), w[hp + "layernorm.weight"], w[hp + "layernorm.bias"], 1e-6)
    y = F.linear(F.gelu(F.linear(y, w[hp + "mlp.fc1.weight"], w[hp + "mlp.fc1.bias"]), approximate="tanh"),
                 w[hp + "mlp.fc2.weight"], w[hp + "mlp.fc2.bias"])
    return (x + y)[:, 0]


def legacy_siglip_vision_encode(w, arch, frames_0_255, frame_token_pooled=(7, 7), frame_token_cls=False):
    """models/vision_live.py:11-31: normalize(frames/255, .5, .5) -> all `vit_layers_total` layers -> post_layernorm ->
    adaptive_avg_pool2d spatial tokens, optionally preceded by the CLS token (pooler_output)."""
    x = (frames_0_255 * 0.00392156862745098 - 0.5) / 0.5
    h = siglip_tower(w, arch, x, n_layers=arch.vit_layers_total)
    h = F.layer_norm(h, (h.shape[-1],), w[VT + "post_layernorm.weight"], w[VT + "post_layernorm.bias"], 1e-6)
    sp = None
    if frame_token_pooled:
        s = int(math.sqrt(h.shape[1]))
        sp = F.adaptive_avg_pool2d(h.reshape(h.shape[0], s, s, h.shape[-1]).permute(0, 3, 1, 2), tuple(frame_token_pooled))
        sp = sp.flatten(2, 3).permute(0, 2, 1)
        if not frame_token_cls:
            return sp
    cls = siglip_pooling_head(w, arch, h)[:, None]
    return cls if sp is None else torch.cat([cls, sp], dim=1)


# ------------------------------------------------------------------------------------------------------------
# Qwen2 decoder with KV append  (TF:models/qwen2/modeling_qwen2.py:35-48 MLP, :102-146 RoPE, :187-246 attention,
# :249-263 RMSNorm, :269-310 layer, :353-409 model; cache TF:cache_utils.py:88-152)
# ------------------------------------------------------------------------------------------------------------
def rms_norm(x, weight, eps):
    dt = x.dtype
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return weight * xf.to(dt)


def rope_cos_sin(arch, positions, dtype):
    inv_freq = 1.0 / (arch.rope_theta ** (torch.arange(0, arch.head_dim, 2, dtype=torch.int64).float() / arch.head_dim))
    freqs = positions.float()[:, None] * inv_freq[None, :].to(positions.device)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


class KVCache:
    """Per-layer [kv_heads, L, dh] tensors that grow by concatenation (DynamicCache.update) and can be cropped."""

    def __init__(self, n_layers):
        self.k = [None] * n_layers
        self.v = [None] * n_layers

    def __len__(self):
        return 0 if self.k[0] is None else self.k[0].shape[1]

    def append(self, i, k, v):
        self.k[i] = k if self.k[i] is None else torch.cat([self.k[i], k], dim=1)
        self.v[i] = v if self.v[i] is None else torch.cat([self.v[i], v], dim=1)
        return self.k[i], self.v[i]

    def crop(self, length):
        for i in range(len(self.k)):
            if self.k[i] is not None:
                self.k[i] = self.k[i][:, :length]
                self.v[i] = self.v[i][:, :length]


def decoder_forward(w, arch, inputs_embeds, cache, attn_impl="eager"):
    """Qwen2Model.forward for one stream: inputs_embeds [M, hidden], appends M tokens to `cache`; returns the final
    RMSNorm output [M, hidden].  attn_impl: "eager" (the oracle of record: fp32 softmax, as HF's eager attention) or
    "sdpa" (torch SDPA as the reference's default `attn_implementation`; only bench.py's baseline timing uses it)."""
    M = inputs_embeds.shape[0]
    past = len(cache)
    pos = torch.arange(past, past + M, device=inputs_embeds.device)
    cos, sin = rope_cos_sin(arch, pos, inputs_embeds.dtype)
    Hq, Hkv, dh = arch.q_heads, arch.kv_heads, arch.head_dim
    x = inputs_embeds
    # bottom-right aligned causal mask over [past + M] keys
    mask = torch.ones(M, past + M, dtype=torch.bool, device=x.device).tril(diagonal=past)
    for i in range(arch.layers):
        p = f"model.layers.{i}."
        h = rms_norm(x, w[p + "input_layernorm.weight"], arch.rms_eps)
        q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]).view(M, Hq, dh).transpose(0, 1)
        k = F.linear(h, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"]).view(M, Hkv, dh).transpose(0, 1)
        v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]).view(M, Hkv, dh).transpose(0, 1)
        q = q * cos[None] + rotate_half(q) * sin[None]
        k = k * cos[None] + rotate_half(k) * sin[None]
        kk, vv = cache.append(i, k, v)
        kk = kk.repeat_interleave(Hq // Hkv, dim=0)
        vv = vv.repeat_interleave(Hq // Hkv, dim=0)
        if attn_impl == "sdpa":
            o = F.scaled_dot_product_attention(q[None], kk[None], vv[None], attn_mask=mask[None, None])[0]
            o = o.transpose(0, 1).reshape(M, Hq * dh)
        else:
            s = (q @ kk.transpose(-1, -2)) * dh ** -0.5
            s = s.masked_fill(~mask[None], float("-inf"))
            a = torch.softmax(s.float(), dim=-1).to(q.dtype)
            o = (a @ vv).transpose(0, 1).reshape(M, Hq * dh)
        x = x + F.linear(o, w[p + "self_attn.o_proj.weight"])
        h = rms_norm(x, w[p + "post_attention_layernorm.weight"], arch.rms_eps)
        h = F.silu(F.linear(h, w[p + "mlp.gate_proj.weight"])) * F.linear(h, w[p + "mlp.up_proj.weight"])
        x = x + F.linear(h, w[p + "mlp.down_proj.weight"])
    return rms_norm(x, w["model.norm.weight"], arch.rms_eps)


def model_forward(w, arch, inputs_embeds, cache, want_lm_logits=False, attn_impl="eager"):
    """VideoHeadLiveLlavaQwenForCausalLM.forward (video_head_live_llava_qwen.py:121-205) on inputs_embeds [M, hidden]:
    returns dict(informative_logits [M,2] fp32, relevance_logits [M,2] fp32, logits [M,V] fp32 (optional))."""
    h = decoder_forward(w, arch, inputs_embeds, cache, attn_impl)
    out = {
        "informative_logits": F.linear(h, w["informative_head.weight"]).float(),
        "relevance_logits": F.linear(h, w["relevance_head.weight"]).float(),
        "hidden": h,
    }
    if want_lm_logits:
        out["logits"] = F.linear(h, w["lm_head.weight"]).float()
    return out


def embed_tokens(w, ids):
    return F.embedding(ids, w["model.embed_tokens.weight"])


# ------------------------------------------------------------------------------------------------------------
# frame loop  (test/inference.py:196-313, demo/liveinfer.py:69-105) — host logic; generation stubbed by callback
# ------------------------------------------------------------------------------------------------------------
class LiveLoopOracle:
    """Restates LiveInferForBenchmark's per-frame control flow over the functional model above.

    Tokenizer-dependent pieces are parameters (the tokenizer files are not available offline): `start_ids`
    (system prompt, test/inference.py:61), `stream_prompt_ids` (:62), `stream_generation_ids` (:63).  Generation
    (`_generate_response`, :257-274) runs the greedy loop of models/modeling_live.py:51-77 with `max_new_tokens`."""

    def __init__(self, w, arch, *, start_ids, stream_prompt_ids=(), stream_generation_ids=(), eos_token_id=None,
                 frame_fps=2.0, score_heads=("informative_score",), stream_end_prob_threshold=None,
                 stream_end_score_sum_threshold=None, running_list_length=20, remove_assistant_turns=False,
                 max_new_tokens=200, dtype=torch.float32, repetition_penalty=None):
        if int(stream_end_prob_threshold is not None) + int(stream_end_score_sum_threshold is not None) != 1:
            raise ValueError("only one of stream_end_prob_threshold / stream_end_score_sum_threshold can be set")
        self.w, self.arch, self.dtype = w, arch, dtype
        self.start_ids = torch.as_tensor(start_ids, dtype=torch.long)
        self.stream_prompt_ids = torch.as_tensor(stream_prompt_ids, dtype=torch.long)
        self.stream_generation_ids = torch.as_tensor(stream_generation_ids, dtype=torch.long)
        self.eos_token_id = eos_token_id
        self.frame_fps = frame_fps
        self.score_heads = list(score_heads)
        self.stream_end_prob_threshold = stream_end_prob_threshold
        self.stream_end_score_sum_threshold = stream_end_score_sum_threshold
        self.running_list_length = running_list_length
        self.remove_assistant_turns = remove_assistant_turns
        self.max_new_tokens = max_new_tokens
        self.repetition_penalty = repetition_penalty
        self.top2_gaps = []        # (kind, gap between the two largest lm logits) of every argmax the loop takes
        self.reset()

    def reset(self):
        import collections
        self.query_queue = collections.deque()
        self.frame_embeds_queue = collections.deque()
        self.video_time = 0
        self.frame_idx = 0
        self.last_role = "system"
        self.last_ids = torch.zeros(0, dtype=torch.long)
        self.cache = KVCache(self.arch.layers)
        self.debug_data_list = []
        self.stream_end_prob_list = []
        self.stream_end_score_sum = 0
        self.generated_token_ids = []

    def input_video_stream(self, pixels, batch_size=32):
        for b in range(0, len(pixels), batch_size):
            emb = visual_embed(self.w, self.arch, pixels[b:b + batch_size].to(self.dtype)).split(self.arch.frame_tokens)
            self.frame_embeds_queue.extend([((r + b) / self.frame_fps, f) for r, f in enumerate(emb)])

    def input_query_stream(self, queries):
        """queries: iterable of (time, token_ids) — already chat-templated user turns — or (time, f) with f(last_role) -> ids
        (the template's stream-query prefix depends on the role of the turn before the query, test/inference.py:250)."""
        for t, ids in queries:
            self.query_queue.append((t, ids if callable(ids) else torch.as_tensor(ids, dtype=torch.long)))

    def _encode_frame(self):
        _, frame_embeds = self.frame_embeds_queue.popleft()
        if len(self.cache) == 0:
            self.last_ids = self.start_ids
        elif self.last_role == "assistant" and not self.remove_assistant_turns:
            self.last_ids = torch.cat([self.last_ids, self.stream_prompt_ids])
        else:
            self.last_ids = torch.zeros(0, dtype=torch.long)
        x = torch.cat([embed_tokens(self.w, self.last_ids).to(self.dtype), frame_embeds.to(self.dtype)], dim=0)
        out = model_forward(self.w, self.arch, x, self.cache)
        self.frame_idx += 1
        self.last_role = "stream"
        return {"informative_score": out["informative_logits"][-1].softmax(-1)[1].item(),
                "relevance_score": out["relevance_logits"][-1].softmax(-1)[1].item()}

    def _encode_query(self):
        _, ids = self.query_queue.popleft()
        if callable(ids):
            ids = torch.as_tensor(ids(self.last_role), dtype=torch.long)
        out = model_forward(self.w, self.arch, embed_tokens(self.w, ids).to(self.dtype), self.cache, want_lm_logits=True)
        self.last_ids = out["logits"][-1:].argmax(-1)
        self._note_gap("query", out["logits"][-1])
        self.last_role = "user"

    def _note_gap(self, kind, logits):
        t = logits.float().topk(2).values
        self.top2_gaps.append((kind, float(t[0] - t[1])))

    def _generate_response(self):
        L0 = len(self.cache)
        x = embed_tokens(self.w, self.stream_generation_ids).to(self.dtype)
        ids = []
        for _ in range(self.max_new_tokens):
            out = model_forward(self.w, self.arch, x, self.cache, want_lm_logits=True)
            logits = out["logits"][-1].clone()
            if self.repetition_penalty is not None and self.generated_token_ids:
                # transformers.RepetitionPenaltyLogitsProcessor on the ids generated so far in this video
                # (models/modeling_live.py:58-66): score < 0 -> score * p, else score / p
                idx = torch.tensor(self.generated_token_ids, dtype=torch.long, device=logits.device)
                sc = logits[idx]
                logits[idx] = torch.where(sc < 0, sc * self.repetition_penalty, sc / self.repetition_penalty)
            tok = logits[None].argmax(-1)
            self._note_gap("generate", logits)
            ids.append(int(tok))
            if self.repetition_penalty is not None and int(tok) != self.eos_token_id:   # modeling_live.py:68-69
                self.generated_token_ids.append(int(tok))
            if self.eos_token_id is not None and int(tok) == self.eos_token_id:
                break
            x = embed_tokens(self.w, tok).to(self.dtype)
        if self.remove_assistant_turns:
            self.cache.crop(L0)  # transformers 4.44.2 meaning: the returned cache is dropped
            self.last_ids = torch.zeros(0, dtype=torch.long)
        else:
            self.last_ids = torch.tensor(ids[-1:], dtype=torch.long)
        self.last_role = "assistant"
        return ids

    def inference(self):
        responses = []
        while self.frame_embeds_queue:
            if self.query_queue and self.video_time >= self.query_queue[0][0]:
                self._encode_query()
            scores = self._encode_frame()
            self.debug_data_list.append(dict(time=self.video_time, **scores))
            need = False
            s = sum(v for k, v in scores.items() if k in self.score_heads)
            self.stream_end_prob_list.append(s)
            self.stream_end_score_sum += s
            if isinstance(self.running_list_length, int) and self.running_list_length > 0:
                self.stream_end_prob_list = self.stream_end_prob_list[-self.running_list_length:]
            if self.stream_end_score_sum_threshold is not None and self.stream_end_score_sum > self.stream_end_score_sum_threshold:
                need = True
                self.stream_end_score_sum = 0
            if self.stream_end_prob_threshold is not None and s > self.stream_end_prob_threshold:
                need = True
            if need:
                responses.append({"time": self.video_time, "content": self._generate_response(), "role": "assistant"})
            self.video_time += 1 / self.frame_fps
        return responses


# ------------------------------------------------------------------------------------------------------------
# deterministic weights and synthetic frames (shared by the oracle, the fixtures and the CUDA path)
# ------------------------------------------------------------------------------------------------------------
def make_weights(arch, seed=1234, device="cpu", dtype=torch.float32, round_bf16=True, include_lm_head=True,
                 legacy_post_ln=False, generate_on_device=False):
    """Random-init state_dict with the reference's key names.  Scales follow the HF / torch default initialisers
    (SigLIP: lecun-normal patch embedding, xavier attention/MLP; mm_projector: nn.Linear default, std 1/sqrt(3*fan_in);
    Qwen2: normal(0, 0.02)), because the north-star tolerance (max-abs 2e-2) is absolute and was stated at that scale
    (frame embeddings abs-max ~4.6, rms ~1.1).  Norm scales/biases get small perturbations so
    that every scale/bias path is exercised.  Values are rounded to bf16 when `round_bf16` so that the fp32 oracle and
    the bf16 kernels see identical weights."""
    gen_dev = device if generate_on_device else "cpu"   # full-size weights: seconds on the GPU, minutes on CPU
    g = torch.Generator(device=gen_dev).manual_seed(seed)
    w = {}

    def rnd(*shape, std=0.02):
        t = torch.randn(*shape, generator=g, dtype=torch.float32, device=gen_dev) * std
        if round_bf16:
            t = t.bfloat16().float()
        return t.to(device=device, dtype=dtype)

    D, Dm = arch.vit_dim, arch.vit_mlp
    w[VT + "embeddings.patch_embedding.weight"] = rnd(D, 3, arch.patch_size, arch.patch_size, std=(3 * arch.patch_size ** 2) ** -0.5)
    w[VT + "embeddings.patch_embedding.bias"] = rnd(D, std=0.02)
    w[VT + "embeddings.position_embedding.weight"] = rnd(arch.patches, D, std=D ** -0.5)
    n_vit = arch.vit_layers_total if legacy_post_ln else arch.vit_layers
    for i in range(n_vit):
        p = f"{VT}encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            w[p + ln + ".weight"] = 1.0 + rnd(D, std=0.05)
            w[p + ln + ".bias"] = rnd(D, std=0.05)
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[p + f"self_attn.{proj}.weight"] = rnd(D, D, std=D ** -0.5)
            w[p + f"self_attn.{proj}.bias"] = rnd(D, std=0.02)
        w[p + "mlp.fc1.weight"] = rnd(Dm, D, std=(2.0 / (D + Dm)) ** 0.5)
        w[p + "mlp.fc1.bias"] = rnd(Dm, std=0.02)
        w[p + "mlp.fc2.weight"] = rnd(D, Dm, std=(2.0 / (D + Dm)) ** 0.5)
        w[p + "mlp.fc2.bias"] = rnd(D, std=0.02)
    if legacy_post_ln:
        w[VT + "post_layernorm.weight"] = 1.0 + rnd(D, std=0.05)
        w[VT + "post_layernorm.bias"] = rnd(D, std=0.05)
        # attention-pooling head (frame_token_cls); its own generator so that no other tensor of the dict depends on it
        g_main, g = g, torch.Generator(device=gen_dev).manual_seed(seed + 977)
        hp = VT + "head."
        w[hp + "probe"] = rnd(1, 1, D, std=1.0)
        w[hp + "attention.in_proj_weight"] = rnd(3 * D, D, std=D ** -0.5)
        w[hp + "attention.in_proj_bias"] = rnd(3 * D, std=0.02)
        w[hp + "attention.out_proj.weight"] = rnd(D, D, std=D ** -0.5)
        w[hp + "attention.out_proj.bias"] = rnd(D, std=0.02)
        w[hp + "layernorm.weight"] = 1.0 + rnd(D, std=0.05)
        w[hp + "layernorm.bias"] = rnd(D, std=0.05)
        w[hp + "mlp.fc1.weight"] = rnd(Dm, D, std=(2.0 / (D + Dm)) ** 0.5)
        w[hp + "mlp.fc1.bias"] = rnd(Dm, std=0.02)
        w[hp + "mlp.fc2.weight"] = rnd(D, Dm, std=(2.0 / (D + Dm)) ** 0.5)
        w[hp + "mlp.fc2.bias"] = rnd(D, std=0.02)
        g = g_main
    H = arch.hidden
    w["model.mm_projector.0.weight"] = rnd(H, D, std=(3 * D) ** -0.5)
    w["model.mm_projector.0.bias"] = rnd(H, std=0.02)
    w["model.mm_projector.2.weight"] = rnd(H, H, std=(3 * H) ** -0.5)
    w["model.mm_projector.2.bias"] = rnd(H, std=0.02)
    w["model.embed_tokens.weight"] = rnd(arch.vocab, H, std=0.02)
    kvd = arch.kv_heads * arch.head_dim
    for i in range(arch.layers):
        p = f"model.layers.{i}."
        w[p + "input_layernorm.weight"] = 1.0 + rnd(H, std=0.05)
        w[p + "post_attention_layernorm.weight"] = 1.0 + rnd(H, std=0.05)
        w[p + "self_attn.q_proj.weight"] = rnd(H, H, std=0.02)
        w[p + "self_attn.q_proj.bias"] = rnd(H, std=0.1)
        w[p + "self_attn.k_proj.weight"] = rnd(kvd, H, std=0.02)
        w[p + "self_attn.k_proj.bias"] = rnd(kvd, std=0.1)
        w[p + "self_attn.v_proj.weight"] = rnd(kvd, H, std=0.02)
        w[p + "self_attn.v_proj.bias"] = rnd(kvd, std=0.1)
        w[p + "self_attn.o_proj.weight"] = rnd(H, H, std=0.02)
        w[p + "mlp.gate_proj.weight"] = rnd(arch.mlp, H, std=0.02)
        w[p + "mlp.up_proj.weight"] = rnd(arch.mlp, H, std=0.02)
        w[p + "mlp.down_proj.weight"] = rnd(H, arch.mlp, std=0.02)
    w["model.norm.weight"] = 1.0 + rnd(H, std=0.05)
    if include_lm_head:
        w["lm_head.weight"] = rnd(arch.vocab, H, std=0.02)
    w["informative_head.weight"] = rnd(2, H, std=0.02)
    w["relevance_head.weight"] = rnd(2, H, std=0.02)
    if round_bf16:  # the "1.0 + noise" LayerNorm/RMSNorm scales are sums: round them too
        for k in w:
            w[k] = w[k].float().bfloat16().float().to(dtype)
    return w


def synthetic_frames(n, seed=0, size=384):
    """uint8 [n, 3, size, size]: low-frequency sinusoid fields drifting over time plus 8-bit noise."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, size), torch.linspace(0, 1, size), indexing="ij")
    frames = []
    ph = torch.rand(3, 3, generator=g) * 6.28
    fr = 1.0 + torch.rand(3, 3, generator=g) * 5.0
    for t in range(n):
        chans = []
        for c in range(3):
            f = (torch.sin(6.28 * fr[c, 0] * xx + ph[c, 0] + 0.21 * t) * torch.cos(6.28 * fr[c, 1] * yy + ph[c, 1] - 0.13 * t)
                 + 0.5 * torch.sin(6.28 * fr[c, 2] * (xx + yy) + ph[c, 2] + 0.37 * t))
            chans.append(f)
        img = torch.stack(chans) * 60.0 + 128.0 + torch.randn(3, size, size, generator=g) * 6.0
        frames.append(img.clamp(0, 255).to(torch.uint8))
    return torch.stack(frames)
