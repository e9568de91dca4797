"""Host-side mirror of the reference's model surface for the hot path, backed by the CUDA engines.

Mirrors (same names, argument meaning and error behaviour):
  * LiveMixin.visual_embed / joint_embed                         models/modeling_live.py:26-48
  * fast_greedy_generate                                         models/modeling_live.py:51-77
  * VideoHeadLiveLlavaQwenForCausalLM.forward / output type      models/live_llava/video_head_live_llava_qwen.py:48-58,121-205
  * build_live_vision / _siglip_vision_encode (legacy entry)     models/vision_live.py:11-31,57-64
PyTorch provides tensors, streams and indexing; every arithmetic step runs in libmmduet_b200.so."""
from dataclasses import dataclass
from typing import Any, Optional

import torch

from . import _lib
from .config import ModelConfig
from .engine import CacheView, DecoderEngine, VisionEngine


@dataclass
class VideoHeadCausalLMOutputWithPast:
    """Field-for-field the reference's output dataclass (video_head_live_llava_qwen.py:48-58)."""
    loss: Optional[Any] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[Any] = None
    hidden_states: Optional[Any] = None
    attentions: Optional[Any] = None
    lm_loss: Optional[Any] = None
    video_loss: Optional[Any] = None
    informative_logits: Optional[torch.Tensor] = None
    relevance_logits: Optional[torch.Tensor] = None


class SigLipImageProcessor:
    """LLaVA's SigLipImageProcessor for frames that are already frame_resolution x frame_resolution (test/datasets.py pads
    to 384x384): rescale 1/255 then normalise with mean = std = 0.5.  The arithmetic itself is fused into the patch
    im2col kernel when uint8 frames are handed to visual_embed(normalize=True); this class exists for API parity
    (test/inference.py:27,203) and returns float pixel_values."""

    image_mean = (0.5, 0.5, 0.5)
    image_std = (0.5, 0.5, 0.5)
    rescale_factor = 0.00392156862745098

    def __init__(self, size=384):
        self.size = (size, size)

    def preprocess(self, images, return_tensors="pt"):
        if not torch.is_tensor(images):
            images = torch.stack([torch.as_tensor(i) for i in images])
        if tuple(images.shape[-2:]) != self.size:
            raise ValueError(f"frames must already be {self.size[0]}x{self.size[1]} (got {tuple(images.shape[-2:])})")
        x = (images.float() * self.rescale_factor - 0.5) / 0.5
        return {"pixel_values": x}


class _VisionTowerHandle:
    def __init__(self, cfg):
        self.image_processor = SigLipImageProcessor(cfg.image_size)
        self.num_patches_per_side = cfg.grid


class _Config:
    """The config attributes the frame loop reads (test/inference.py:33-38,60; video_head_live_llava_qwen.py:101,107)."""

    def __init__(self, cfg: ModelConfig, **extra):
        self.hidden_size = cfg.hidden
        self.vocab_size = cfg.vocab
        self.frame_resolution = cfg.image_size
        self.frame_num_tokens = cfg.frame_tokens
        self.video_pooling_stride = cfg.pool_stride
        self.mm_spatial_pool_mode = cfg.pool_mode
        self.v_placeholder = "<image>"
        self.v_placeholder_id = None
        self.eos_token_id = None
        for k, v in extra.items():
            setattr(self, k, v)


class _Embedding:
    def __init__(self, table):
        self.weight = table

    def __call__(self, ids):
        return self.weight[ids]  # row gather (plumbing); the decoder step gathers ids itself on the fast path


class VideoHeadLiveLlavaQwenForCausalLM:
    """Drop-in for the reference model object on the inference path."""

    def __init__(self, cfg: ModelConfig, state_dict, device="cuda", max_context=None, kv_pages=None, max_step_tokens=512,
                 **config_extra):
        if cfg.pool_mode not in ("bilinear", "average", "max"):
            raise ValueError(f"Unexpected mm_spatial_pool_mode: {cfg.pool_mode}")
        self.model_config = cfg
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = torch.bfloat16
        self.vision = VisionEngine(cfg, state_dict, self.device)
        self.decoder = DecoderEngine(cfg, state_dict, self.device, n_pages=kv_pages, max_tokens=max_step_tokens,
                                     max_context=max_context)
        self.config = _Config(cfg, **config_extra)
        self.vocab_size = cfg.vocab
        self._tower = _VisionTowerHandle(cfg)
        self._embed = _Embedding(self.decoder.embed)
        self.vision_encoder = self._tower  # attribute the reference sets in __init__ (video_head_live_llava_qwen.py:82)

    # ---- trivial nn.Module-like surface ----
    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        return self

    def to(self, *a, **k):
        return self

    def get_model(self):
        return self

    def get_vision_tower(self):
        return self._tower

    def get_input_embeddings(self):
        return self._embed

    def set_vision_inside(self):
        pass  # the vision encoder always lives inside (modeling_live.py:14-20 takes the 'already exists' branch)

    # ---- LiveMixin ----
    @torch.no_grad()
    def visual_embed(self, frames):
        """frames: [T,3,384,384] image-processed pixel_values (any float dtype) -> [T*tokens, hidden] bf16.
        uint8 frames are accepted too and are rescaled/normalised inside the patch-embed loader."""
        if frames.dtype == torch.uint8:
            return self.vision.visual_embed(frames.to(self.device), normalize=True)
        if frames.dtype not in (torch.bfloat16, torch.float32):
            frames = frames.float()
        return self.vision.visual_embed(frames.to(self.device))

    @torch.no_grad()
    def joint_embed(self, input_ids=None, frames=None):
        if frames is None:
            return self.get_input_embeddings()(input_ids)
        if input_ids is None:
            return self.visual_embed(frames)
        inputs_embeds = self.get_input_embeddings()(input_ids.clamp(max=self.vocab_size - 1)).clone()
        v_mask = input_ids == self.config.v_placeholder_id
        if v_mask.any():
            inputs_embeds[v_mask] = self.visual_embed(frames).to(inputs_embeds.dtype)
        return inputs_embeds

    # ---- forward ----
    def new_cache(self):
        return CacheView(self.decoder.new_stream(), 0)

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, informative_labels=None, relevance_labels=None, use_cache=None, output_attentions=None,
                output_hidden_states=None, frames=None, return_dict=None, logits_to_keep="all", **kwargs):
        """Batch-1 forward over `inputs_embeds` [1,M,H] with KV append.  `logits_to_keep`: 'all' (reference behaviour:
        lm_head on every position), 'last' or 'none' (frame steps never read the lm logits)."""
        if labels is not None or informative_labels is not None or relevance_labels is not None:
            raise NotImplementedError("training losses are outside the accelerated path")
        if inputs_embeds is None:
            inputs_embeds = self.joint_embed(input_ids, frames)
        if inputs_embeds.dim() == 2:
            inputs_embeds = inputs_embeds[None]
        if inputs_embeds.shape[0] != 1:
            raise ValueError("the streaming path is batch-1 per stream (use DecoderEngine.step for several streams)")
        if not past_key_values:
            past_key_values = past_key_values if isinstance(past_key_values, CacheView) else self.new_cache()
        view = past_key_values
        M = inputs_embeds.shape[1]
        lm = {"all": "all", "last": "last", "none": "none"}[logits_to_keep]
        out = self.decoder.step([dict(storage=view.storage, past=view.length, embeds=inputs_embeds[0])], score="all", lm=lm)
        hl = out["head_logits"]
        logits = None
        if lm == "all":
            logits = out["lm_logits"][None]
        elif lm == "last":
            logits = out["lm_logits"][None]  # [1,1,V]: indexable as logits[:, -1:]
        return VideoHeadCausalLMOutputWithPast(
            loss=0.0, logits=logits, past_key_values=out["views"][0], lm_loss=0.0, video_loss=0.0,
            informative_logits=hl[None, :, 0:2], relevance_logits=hl[None, :, 2:4])

    __call__ = forward


def fast_greedy_generate(*, model, inputs_embeds, past_key_values, eos_token_id, inplace_output_ids, repetition_penalty=None,
                         generated_token_ids=list()):
    """models/modeling_live.py:51-77: token-by-token greedy decode (M=1 steps over the paged KV, lm_head on the last row
    only, argmax + HF repetition penalty on the device)."""
    if repetition_penalty is not None:
        assert isinstance(repetition_penalty, float)
    lib = _lib.load()
    dec = model.decoder
    view = past_key_values
    x = inputs_embeds[0] if inputs_embeds.dim() == 3 else inputs_embeds
    new_id = torch.zeros(1, dtype=torch.long, device=model.device)
    # penalised ids live on the device and grow by one device-to-device copy per token (no per-token H2D upload)
    n_pen, pen_buf = 0, None
    if repetition_penalty is not None:
        n_pen = len(generated_token_ids)
        pen_buf = torch.zeros(n_pen + inplace_output_ids.size(1) + 1, dtype=torch.long, device=model.device)
        if n_pen:
            pen_buf[:n_pen] = torch.tensor(generated_token_ids, dtype=torch.long, device=model.device)
    i = 0
    for i in range(inplace_output_ids.size(1)):
        out = dec.step([dict(storage=view.storage, past=view.length, embeds=x)], score="none", lm="last")
        view = out["views"][0]
        rc = lib.mmd_argmax(out["lm_logits"].data_ptr(), model.vocab_size, _lib.ptr(pen_buf) if n_pen else 0, n_pen,
                            float(repetition_penalty or 1.0), new_id.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "mmd_argmax")
        tok = int(new_id.item())
        if repetition_penalty is not None and tok != eos_token_id:
            generated_token_ids.append(tok)
            pen_buf[n_pen:n_pen + 1].copy_(new_id)
            n_pen += 1
        inplace_output_ids[:, i] = tok
        if tok == eos_token_id:
            break
        x = model.get_input_embeddings()(new_id)
    return inplace_output_ids[:, :i + 1], view, generated_token_ids


# ------------------------------------------------------------------------------------------------------------
# legacy entry (models/vision_live.py)
# ------------------------------------------------------------------------------------------------------------
def _siglip_vision_encode(vision_model: VisionEngine, frames, frame_token_cls: bool = False, frame_token_pooled=(7, 7), **kwargs):
    """models/vision_live.py:11-31: frames are raw 0..255 values (float or uint8) -> [T, (1 +) ph*pw, D]: the optional CLS
    token (pooler_output of the SigLIP attention-pooling head) followed by the adaptive-avg-pooled spatial tokens."""
    return vision_model.legacy_encode(frames, tuple(frame_token_pooled) if frame_token_pooled else None, frame_token_cls=bool(frame_token_cls))


def build_live_vision(config, state_dict=None, device="cuda"):
    """models/vision_live.py:57-64.  `config` needs vision_pretrained / frame_token_cls / frame_token_pooled; the weights
    come from `state_dict` (HF SiglipVisionModel keys under model.vision_tower.vision_tower.vision_model.*)."""
    from functools import partial
    name = getattr(config, "vision_pretrained", None)
    if name not in ("google/siglip-large-patch16-384", "google/siglip-so400m-patch14-384"):
        raise ValueError(f"Unverified vision_pretrained: {name}")
    cfg = getattr(config, "model_config", None) or ModelConfig()
    eng = VisionEngine(cfg, state_dict, device, with_projector=False, n_layers=cfg.vit_layers_total, legacy_post_ln=True)
    return eng, partial(_siglip_vision_encode, frame_token_cls=config.frame_token_cls, frame_token_pooled=config.frame_token_pooled)
