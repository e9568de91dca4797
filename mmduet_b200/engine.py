"""Host-side runtime over the C ABI: weight packing (reference state_dict keys -> kernel layouts), the vision encoder
(SigLIP tower + projector + pooling) and the decoder with its paged KV pool.  PyTorch is used for device memory,
streams and host<->device copies only; all arithmetic on the path happens in libmmduet_b200.so."""
import ctypes
import os
import math
import threading

import numpy as np
import torch

from . import _lib
from .config import ModelConfig

VT = "model.vision_tower.vision_tower.vision_model."
PAGE = _lib.PAGE_TOKENS


def _bf16(t, device):
    return t.detach().to(device=device, dtype=torch.bfloat16).contiguous()


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def pooling_taps(grid, stride, mode, out_size=None):
    """Tap matrix [n_out, grid*grid] of the reference's pooling, obtained by pushing an identity basis through torch's
    own operators (video_head_live_llava_qwen.py:107-114; vision_live.py:19-25 for 'adaptive')."""
    import torch.nn.functional as F
    S = grid * grid
    eye = torch.eye(S, dtype=torch.float32).view(1, S, grid, grid)
    if mode == "bilinear":
        out = F.interpolate(eye, size=[math.ceil(grid / stride)] * 2, mode="bilinear")
    elif mode == "average":
        out = F.avg_pool2d(eye, stride)
    elif mode == "max":
        out = F.max_pool2d(eye, stride)
    elif mode == "adaptive":
        out = F.adaptive_avg_pool2d(eye, out_size)
    else:
        raise ValueError(f"Unexpected mm_spatial_pool_mode: {mode}")
    return out.view(S, -1).t().contiguous()


def taps_to_tables(taps):
    """[n_out, S] tap matrix -> (gather_idx [G], tap_idx [n_out, max_taps] into the gathered set, tap_w)."""
    used = (taps != 0).any(dim=0).nonzero().flatten()
    remap = {int(s): i for i, s in enumerate(used.tolist())}
    rows = [[(remap[int(j)], float(taps[o, j])) for j in (taps[o] != 0).nonzero().flatten().tolist()] for o in range(taps.shape[0])]
    max_taps = max(len(r) for r in rows)
    idx = np.full((taps.shape[0], max_taps), -1, dtype=np.int32)
    wgt = np.zeros((taps.shape[0], max_taps), dtype=np.float32)
    for o, r in enumerate(rows):
        for j, (i, w) in enumerate(r):
            idx[o, j] = i
            wgt[o, j] = w
    return used.to(torch.int32), torch.from_numpy(idx), torch.from_numpy(wgt), max_taps


class VisionEngine:
    """model.visual_embed(frames) of the reference (models/modeling_live.py:26-33): SigLIP tower (26 layers, pre-post-LN)
    -> mm_projector -> spatial pooling -> [T * tokens, hidden] bf16; also the legacy models/vision_live.py entry."""

    # Frames per encoder launch sequence.  The reference encodes 32 at a time (test/inference.py:208); frames are independent,
    # so the batch size does not change any value.  40 divides the usual 120/200/400-frame videos evenly and fills the
    # attention grid better (40*16*3 CTAs = 12.97 waves of 148): +3.4 % frames/s over 32 (measured, bench.py).
    MAX_BATCH = int(os.environ.get("MMD_ENCODER_BATCH", "40"))

    def __init__(self, cfg: ModelConfig, state_dict, device, with_projector=True, n_layers=None, legacy_post_ln=False,
                 attn_out_split=True, projector_hilo=True):
        cfg.validate()
        self.cfg, self.device = cfg, torch.device(device)
        self.lib = _lib.load()
        self.ctx = _lib.context(self.device.index)
        sd, dev = state_dict, self.device
        D, P = cfg.vit_dim, cfg.patch_size
        k_real = 3 * P * P
        self.k_pad = (k_real + 7) // 8 * 8
        keep = []
        pw = torch.zeros(D, self.k_pad, dtype=torch.bfloat16, device=dev)
        pw[:, :k_real] = _bf16(sd[VT + "embeddings.patch_embedding.weight"].reshape(D, k_real), dev)
        patch_b = _f32(sd[VT + "embeddings.patch_embedding.bias"], dev)
        # the position embedding is an fp32 add on the residual stream; keep the (bf16-representable) parameter in fp32
        pos = _f32(sd[VT + "embeddings.position_embedding.weight"], dev)
        keep += [pw, patch_b, pos]
        self.n_layers = cfg.vit_layers if n_layers is None else n_layers
        layers = (_lib.VitLayer * self.n_layers)()
        for i in range(self.n_layers):
            p = f"{VT}encoder.layers.{i}."
            qkv_w = _bf16(torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0), dev)
            qkv_b = _f32(torch.cat([sd[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0), dev)
            t = dict(ln1_w=_f32(sd[p + "layer_norm1.weight"], dev), ln1_b=_f32(sd[p + "layer_norm1.bias"], dev),
                     qkv_w=qkv_w, qkv_b=qkv_b,
                     out_w=(_bf16(torch.cat([sd[p + "self_attn.out_proj.weight"]] * 2, 1), dev) if attn_out_split
                            else _bf16(sd[p + "self_attn.out_proj.weight"], dev)),
                     out_b=_f32(sd[p + "self_attn.out_proj.bias"], dev),
                     ln2_w=_f32(sd[p + "layer_norm2.weight"], dev), ln2_b=_f32(sd[p + "layer_norm2.bias"], dev),
                     fc1_w=_bf16(sd[p + "mlp.fc1.weight"], dev), fc1_b=_f32(sd[p + "mlp.fc1.bias"], dev),
                     fc2_w=_bf16(sd[p + "mlp.fc2.weight"], dev), fc2_b=_f32(sd[p + "mlp.fc2.bias"], dev))
            for k, v in t.items():
                setattr(layers[i], k, v.data_ptr())
            keep.append(t)
        self._layers = layers
        self.vit = _lib.VitWeights(image_size=cfg.image_size, patch_size=P, dim=D, heads=cfg.vit_heads, mlp=cfg.vit_mlp,
                                   n_layers=self.n_layers, k_pad=self.k_pad, attn_out_split=int(attn_out_split), patch_w=pw.data_ptr(), patch_b=patch_b.data_ptr(),
                                   pos_emb=pos.data_ptr(), layers=layers)
        self.post_ln = None
        self.cls_head = None
        if legacy_post_ln:
            self.post_ln = (_f32(sd[VT + "post_layernorm.weight"], dev), _f32(sd[VT + "post_layernorm.bias"], dev))
            if VT + "head.probe" in sd:
                # SiglipMultiheadAttentionPoolingHead (frame_token_cls): the probe's query is a constant of the weights
                hp = VT + "head."
                w_in, b_in = sd[hp + "attention.in_proj_weight"].float(), sd[hp + "attention.in_proj_bias"].float()
                q = (sd[hp + "probe"].float().view(1, D) @ w_in[:D].t() + b_in[:D]) * (D // cfg.vit_heads) ** -0.5
                self.cls_head = dict(q=_f32(q.view(-1), dev), kv_w=_bf16(w_in[D:], dev), kv_b=_f32(b_in[D:], dev),
                                     out_w=_bf16(sd[hp + "attention.out_proj.weight"], dev), out_b=_f32(sd[hp + "attention.out_proj.bias"], dev),
                                     ln_w=_f32(sd[hp + "layernorm.weight"], dev), ln_b=_f32(sd[hp + "layernorm.bias"], dev),
                                     fc1_w=_bf16(sd[hp + "mlp.fc1.weight"], dev), fc1_b=_f32(sd[hp + "mlp.fc1.bias"], dev),
                                     fc2_w=_bf16(sd[hp + "mlp.fc2.weight"], dev), fc2_b=_f32(sd[hp + "mlp.fc2.bias"], dev))
        self.proj = None
        if with_projector:
            taps = pooling_taps(cfg.grid, cfg.pool_stride, cfg.pool_mode)
            gidx, tidx, tw, max_taps = taps_to_tables(taps)
            self.tokens_per_frame = taps.shape[0]
            rep = 2 if projector_hilo else 1
            t = dict(w1=_bf16(torch.cat([sd["model.mm_projector.0.weight"]] * rep, 1), dev), b1=_f32(sd["model.mm_projector.0.bias"], dev),
                     w2=_bf16(torch.cat([sd["model.mm_projector.2.weight"]] * rep, 1), dev), b2=_f32(sd["model.mm_projector.2.bias"], dev),
                     gather_idx=gidx.to(dev), tap_idx=tidx.to(dev).contiguous(), tap_w=tw.to(dev).contiguous())
            # Linear pooling (bilinear / average; tap weights sum to 1) commutes with Linear2, so it runs in Linear1's epilogue:
            # tap-major gather, `pool_group` (4 or 16) consecutive rows per output token, weight-0 padding slots
            pool_group = 0
            if projector_hilo and cfg.pool_mode != "max" and max_taps <= 16 and os.environ.get("MMD_PROJ_GENERIC") != "1":
                pool_group = 4 if max_taps <= 4 else 16
                src = torch.zeros(taps.shape[0], pool_group, dtype=torch.int32)
                wgt = torch.zeros(taps.shape[0], pool_group, dtype=torch.float32)
                for o in range(taps.shape[0]):
                    nz = (taps[o] != 0).nonzero().flatten()
                    src[o, :len(nz)] = nz.to(torch.int32)
                    src[o, len(nz):] = int(nz[0])
                    wgt[o, :len(nz)] = taps[o, nz]
                t["pool_gather_idx"], t["pool_row_w"] = src.flatten().to(dev), wgt.flatten().to(dev)
            keep.append(t)
            self.proj = _lib.ProjectorWeights(vit_dim=D, hidden=cfg.hidden, n_src_tokens=cfg.patches, n_gather=int(gidx.numel()),
                                              n_out=self.tokens_per_frame, max_taps=max_taps, maxpool=int(cfg.pool_mode == "max"),
                                              hilo=int(projector_hilo), pool_group=pool_group,
                                              **{k: v.data_ptr() for k, v in t.items()})
        self._keep = keep
        self._ws = {}
        self._legacy_taps = {}

    def _workspace(self, T):
        ws = self._ws.get(T)
        if ws is None:
            n = self.lib.mmd_vit_workspace_bytes(ctypes.byref(self.vit), T)
            if self.proj is not None:
                n = max(n, self.lib.mmd_projector_workspace_bytes(ctypes.byref(self.proj), T))
            ws = (torch.empty(n, dtype=torch.uint8, device=self.device),
                  torch.empty(T * self.cfg.patches, self.cfg.vit_dim, dtype=torch.float32, device=self.device))
            self._ws = {T: ws}  # keep only the latest size
        return ws

    def _pixel_dtype(self, frames):
        if frames.dtype == torch.uint8:
            return _lib.DT_U8
        if frames.dtype == torch.bfloat16:
            return _lib.DT_BF16
        if frames.dtype == torch.float32:
            return _lib.DT_F32
        raise TypeError(f"frames dtype {frames.dtype} unsupported (uint8, bfloat16, float32)")

    def tower(self, frames, normalize=False):
        """[T,3,H,W] pixels -> fp32 hidden state [T*patches, vit_dim] of the last executed layer (pre-post_layernorm).
        The returned tensor is a workspace view that the next call overwrites."""
        assert frames.is_cuda and frames.dim() == 4 and frames.shape[1] == 3, frames.shape
        assert frames.shape[2] == self.cfg.image_size and frames.shape[3] == self.cfg.image_size, frames.shape
        frames = frames.contiguous()
        T = frames.shape[0]
        ws, resid = self._workspace(T)
        rc = self.lib.mmd_vit_forward(self.ctx, ctypes.byref(self.vit), frames.data_ptr(), self._pixel_dtype(frames), int(normalize),
                                      T, resid.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "mmd_vit_forward")
        return resid

    def visual_embed(self, frames, normalize=False, out_dtype=torch.bfloat16, out=None):
        """frames [T,3,384,384] (already image-processed unless normalize=True) -> [T*tokens_per_frame, hidden] in the model
        dtype (bf16, what the reference returns) or in fp32 (the same values before the final rounding).
        `out`: optional preallocated destination [T*tokens_per_frame, hidden]; it may live in ANOTHER GPU's memory
        (a symmetric-memory peer mapping, parallel.PeerStoreEncoder): the pooling epilogue then stores over NVLink."""
        assert out_dtype in (torch.bfloat16, torch.float32)
        if self.proj is None:
            raise _lib.MmdError("this VisionEngine was built without the projector")
        tpf, T_all = self.tokens_per_frame, frames.shape[0]
        if out is None:
            out = torch.empty(T_all * tpf, self.cfg.hidden, dtype=out_dtype, device=self.device)
        elif out.shape != (T_all * tpf, self.cfg.hidden) or out.dtype != out_dtype or not out.is_contiguous():
            raise _lib.MmdError(f"visual_embed: out must be a contiguous [{T_all * tpf}, {self.cfg.hidden}] {out_dtype} tensor")
        for b in range(0, T_all, self.MAX_BATCH):
            chunk = frames[b:b + self.MAX_BATCH]
            T = chunk.shape[0]
            resid = self.tower(chunk, normalize)
            ws, _ = self._workspace(T)
            dst = out[b * tpf:(b + T) * tpf]
            rc = self.lib.mmd_projector_pool(self.ctx, ctypes.byref(self.proj), resid.data_ptr(), T, dst.data_ptr(),
                                             _lib.DT_BF16 if out_dtype == torch.bfloat16 else _lib.DT_F32, ws.data_ptr(), ws.numel(),
                                             _lib.stream_ptr())
            _lib.check(rc, "mmd_projector_pool")
        return out

    def _cls_tokens(self, normed, T):
        """pooler_output of the SigLIP vision model for T frames: attention pooling over the post-layernormed patch tokens
        (probe attention on CUDA cores, every linear layer through the tcgen05 GEMM), then residual MLP.  -> fp32 [T, D]."""
        from . import ops
        h, D, S = self.cls_head, self.cfg.vit_dim, self.cfg.patches
        x = normed.to(torch.bfloat16)                                                      # dtype cast (plumbing)
        kv = ops.gemm(x, h["kv_w"], bias=h["kv_b"])                                         # [T*S, 2D] bf16
        att = torch.empty(T, D, dtype=torch.float32, device=self.device)
        rc = self.lib.mmd_probe_attention(h["q"].data_ptr(), kv.data_ptr(), att.data_ptr(), T, S, self.cfg.vit_heads, D // self.cfg.vit_heads,
                                          _lib.stream_ptr())
        _lib.check(rc, "mmd_probe_attention")
        res = ops.gemm(att.to(torch.bfloat16), h["out_w"], bias=h["out_b"], epi=_lib.EPI_F32)    # residual [T, D] fp32
        hb = torch.empty(T, D, dtype=torch.bfloat16, device=self.device)
        rc = self.lib.mmd_layernorm(res.data_ptr(), h["ln_w"].data_ptr(), h["ln_b"].data_ptr(), hb.data_ptr(), 0, T, D, 1e-6, _lib.stream_ptr())
        _lib.check(rc, "mmd_layernorm")
        mid = ops.gemm(hb, h["fc1_w"], bias=h["fc1_b"], act=_lib.ACT_GELU_TANH)
        ops.gemm(mid, h["fc2_w"], bias=h["fc2_b"], out=res, epi=_lib.EPI_RESID_F32)
        return res

    def legacy_encode(self, frames_0_255, frame_token_pooled=(7, 7), frame_token_cls=False):
        """models/vision_live.py:11-31 semantics: rescale+normalize, all layers, post_layernorm, adaptive_avg_pool2d to
        `frame_token_pooled` (None/empty: no spatial tokens) and, with frame_token_cls, the pooler_output as the first token."""
        if self.post_ln is None:
            raise _lib.MmdError("legacy_encode needs legacy_post_ln=True (post_layernorm weights)")
        if frame_token_cls and self.cls_head is None:
            raise _lib.MmdError("frame_token_cls needs the vision model's head.* weights (SigLIP attention-pooling head)")
        if not frame_token_pooled and not frame_token_cls:
            raise ValueError("frame_token_pooled must be set when frame_token_cls is False")
        outs = []
        key = tuple(frame_token_pooled) if frame_token_pooled else None
        if key is not None and key not in self._legacy_taps:
            gidx, tidx, tw, max_taps = taps_to_tables(pooling_taps(self.cfg.grid, 0, "adaptive", key))
            assert gidx.numel() == self.cfg.patches
            self._legacy_taps[key] = (tidx.to(self.device).contiguous(), tw.to(self.device).contiguous(), max_taps, tidx.shape[0])
        if key is not None:
            tidx, tw, max_taps, n_out = self._legacy_taps[key]
        D, S = self.cfg.vit_dim, self.cfg.patches
        for b in range(0, frames_0_255.shape[0], self.MAX_BATCH):
            chunk = frames_0_255[b:b + self.MAX_BATCH]
            T = chunk.shape[0]
            resid = self.tower(chunk, normalize=True)
            normed = torch.empty(T * S, D, dtype=torch.float32, device=self.device)
            rc = self.lib.mmd_layernorm(resid.data_ptr(), self.post_ln[0].data_ptr(), self.post_ln[1].data_ptr(), normed.data_ptr(), 1,
                                        T * S, D, 1e-6, _lib.stream_ptr())
            _lib.check(rc, "mmd_layernorm")
            parts = []
            if frame_token_cls:
                parts.append(self._cls_tokens(normed, T)[:, None])
            if key is not None:
                out = torch.empty(T, n_out, D, dtype=torch.float32, device=self.device)
                rc = self.lib.mmd_tap_pool(normed.data_ptr(), _lib.DT_F32, out.data_ptr(), _lib.DT_F32, tidx.data_ptr(), tw.data_ptr(), T, S,
                                           n_out, max_taps, D, 0, _lib.stream_ptr())
                _lib.check(rc, "mmd_tap_pool")
                parts.append(out)
            outs.append(parts[0] if len(parts) == 1 else torch.cat(parts, 1))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)


class KVStorage:
    """Pages of one stream (video) in the pool.  `length` = tokens currently valid."""

    def __init__(self, engine):
        self.engine = engine
        self.pages = []
        self.length = 0

    def ensure(self, new_len):
        need = (new_len + PAGE - 1) // PAGE
        while len(self.pages) < need:
            self.pages.append(self.engine._alloc_page())

    def truncate(self, length):
        """O(1) rollback (the 4.44.2 'drop the returned cache' meaning); pages beyond are recycled."""
        assert 0 <= length <= self.length
        self.length = length
        need = (length + PAGE - 1) // PAGE
        while len(self.pages) > need:
            self.engine._free_page(self.pages.pop())

    def release(self):
        self.truncate(0)

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


class CacheView:
    """What the reference calls `past_key_values`: an immutable (storage, length) pair.  Passing an OLDER view back
    into forward() rolls the stream back to that length, exactly like dropping a newer legacy-tuple cache did."""

    def __init__(self, storage, length):
        self.storage, self.length = storage, length

    def get_seq_length(self, layer_idx=0):
        return self.length

    def __len__(self):
        return self.length

    def __bool__(self):
        return self.length > 0


class DecoderEngine:
    """Qwen2 decoder + informative/relevance heads over a paged KV pool (video_head_live_llava_qwen.py:121-205)."""

    def __init__(self, cfg: ModelConfig, state_dict, device, n_pages=None, max_tokens=512, max_lm_rows=1, max_context=None,
                 layer_range=None):
        """layer_range=(l0, l1): this engine is ONE STAGE of a layer pipeline (parallel.LayerPipeline): it holds decoder layers
        l0 .. l1-1 and their KV pages only; the first stage also holds the embedding table, the last one model.norm, the
        heads and lm_head.  step(..., resid_in=, resid_out=True) hands the fp32 residual stream from stage to stage."""
        cfg.validate()
        self.cfg, self.device = cfg, torch.device(device)
        self.lib = _lib.load()
        self.ctx = _lib.context(self.device.index)
        sd, dev = state_dict, self.device
        H = cfg.hidden
        keep = []
        l0, l1 = layer_range if layer_range is not None else (0, cfg.layers)
        if not (0 <= l0 < l1 <= cfg.layers):
            raise _lib.MmdError(f"layer_range {layer_range} outside 0..{cfg.layers}")
        self.layer_range, self.first_stage, self.last_stage = (l0, l1), l0 == 0, l1 == cfg.layers
        layers = (_lib.DecLayer * (l1 - l0))()
        for i in range(l0, l1):
            p = f"model.layers.{i}."
            t = dict(ln1_w=_f32(sd[p + "input_layernorm.weight"], dev),
                     qkv_w=_bf16(torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0), dev),
                     qkv_b=_f32(torch.cat([sd[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0), dev),
                     o_w=_bf16(sd[p + "self_attn.o_proj.weight"], dev),
                     ln2_w=_f32(sd[p + "post_attention_layernorm.weight"], dev),
                     # gate/up rows interleaved (2j = gate_j, 2j+1 = up_j): one weight stream, SwiGLU pairs adjacent
                     gate_up_w=_bf16(torch.stack([sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]], 1)
                                     .reshape(2 * cfg.mlp, H), dev),
                     down_w=_bf16(sd[p + "mlp.down_proj.weight"], dev))
            for k, v in t.items():
                setattr(layers[i - l0], k, v.data_ptr())
            keep.append(t)
        self.max_context = int(max_context or cfg.max_pos)
        # RoPE tables exactly as Qwen2RotaryEmbedding computes them in fp32 (TF:models/qwen2/modeling_qwen2.py:102-125)
        inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.int64).float() / cfg.head_dim))
        freqs = torch.arange(self.max_context, dtype=torch.float32)[:, None] * inv_freq[None, :]
        t = dict(rope_cos=_f32(freqs.cos(), dev), rope_sin=_f32(freqs.sin(), dev))
        if self.first_stage:
            t["embed"] = _bf16(sd["model.embed_tokens.weight"], dev)
        if self.last_stage:
            t["final_norm_w"] = _f32(sd["model.norm.weight"], dev)
            t["heads_w"] = _f32(torch.cat([sd["informative_head.weight"], sd["relevance_head.weight"]], 0), dev)
        self.embed = t.get("embed")
        lm = sd.get("lm_head.weight") if self.last_stage else None
        self.lm_head = _bf16(lm, dev) if lm is not None else None
        keep.append(t)
        self._layers, self._keep = layers, keep
        ptrs = {k: (t[k].data_ptr() if k in t else 0) for k in ("final_norm_w", "embed", "heads_w", "rope_cos", "rope_sin")}
        self.w = _lib.DecWeights(hidden=H, n_layers=l1 - l0, q_heads=cfg.q_heads, kv_heads=cfg.kv_heads, head_dim=cfg.head_dim,
                                 mlp=cfg.mlp, vocab=cfg.vocab, max_pos=self.max_context, rms_eps=cfg.rms_eps, layers=layers,
                                 lm_head=self.lm_head.data_ptr() if self.lm_head is not None else 0, **ptrs)
        if n_pages is None:
            n_pages = (self.max_context + PAGE - 1) // PAGE + 1
        self.n_pages = n_pages
        self.pool = torch.empty(l1 - l0, n_pages, 2, cfg.kv_heads, PAGE, cfg.head_dim, dtype=torch.bfloat16, device=dev)
        self.kv = _lib.KvPool(pool=self.pool.data_ptr(), layer_stride=self.pool.stride(0), n_pages=n_pages)
        self._free = list(range(n_pages - 1, -1, -1))
        self._lock = threading.Lock()
        self.max_tokens, self.max_lm_rows = max_tokens, max_lm_rows
        self._ws = None
        self._ensure_ws(max_tokens, max_lm_rows)
        self._meta_ring, self._meta_next = [None] * 8, 0

    def stage(self, layer_range, n_pages=None, max_tokens=None):
        """A pipeline-stage view of this (whole) engine: decoder layers l0 .. l1-1 on the SAME packed weights, with its own KV
        pages, workspace and lock (parallel.LayerPipeline runs one stage per GPU; tests chain stages on one GPU)."""
        l0, l1 = layer_range
        if self.layer_range != (0, self.cfg.layers) or not (0 <= l0 < l1 <= self.cfg.layers):
            raise _lib.MmdError(f"stage({layer_range}) needs a whole engine and a range inside 0..{self.cfg.layers}")
        st = object.__new__(DecoderEngine)
        st.cfg, st.device, st.lib, st.ctx = self.cfg, self.device, self.lib, self.ctx
        st.layer_range, st.first_stage, st.last_stage = (l0, l1), l0 == 0, l1 == self.cfg.layers
        layers = (_lib.DecLayer * (l1 - l0))()
        for i in range(l0, l1):
            layers[i - l0] = self._layers[i]
        st._layers, st._keep = layers, self._keep              # the weight tensors stay owned (and alive) through the parent's list
        st.max_context = self.max_context
        st.embed = self.embed if st.first_stage else None
        st.lm_head = self.lm_head if st.last_stage else None
        w = self.w
        st.w = _lib.DecWeights(hidden=w.hidden, n_layers=l1 - l0, q_heads=w.q_heads, kv_heads=w.kv_heads, head_dim=w.head_dim, mlp=w.mlp,
                               vocab=w.vocab, max_pos=w.max_pos, rms_eps=w.rms_eps, layers=layers,
                               final_norm_w=w.final_norm_w if st.last_stage else 0, heads_w=w.heads_w if st.last_stage else 0,
                               lm_head=w.lm_head if st.last_stage else 0, embed=w.embed if st.first_stage else 0,
                               rope_cos=w.rope_cos, rope_sin=w.rope_sin)
        st.n_pages = n_pages if n_pages is not None else self.n_pages
        cfg = self.cfg
        st.pool = torch.empty(l1 - l0, st.n_pages, 2, cfg.kv_heads, PAGE, cfg.head_dim, dtype=torch.bfloat16, device=self.device)
        st.kv = _lib.KvPool(pool=st.pool.data_ptr(), layer_stride=st.pool.stride(0), n_pages=st.n_pages)
        st._free = list(range(st.n_pages - 1, -1, -1))
        st._lock = threading.Lock()
        st.max_tokens, st.max_lm_rows = max_tokens or self.max_tokens, self.max_lm_rows
        st._ws = None
        st._ensure_ws(st.max_tokens, st.max_lm_rows)
        st._meta_ring, st._meta_next = [None] * 8, 0
        return st

    # ---- pages ----
    def _alloc_page(self):
        if not self._free:
            raise _lib.MmdError(f"KV pool exhausted ({self.n_pages} pages of {PAGE} tokens)")
        return self._free.pop()

    def _free_page(self, p):
        self._free.append(p)

    def new_stream(self):
        return KVStorage(self)

    def _ensure_ws(self, n_tokens, n_lm):
        if self._ws is None or n_tokens > self.max_tokens or n_lm > self.max_lm_rows:
            self.max_tokens, self.max_lm_rows = max(self.max_tokens, n_tokens), max(self.max_lm_rows, n_lm)
            n = self.lib.mmd_decoder_workspace_bytes(self.ctx, ctypes.byref(self.w), self.max_tokens, self.max_lm_rows)
            self._ws = torch.empty(n, dtype=torch.uint8, device=self.device)

    # ---- the step ----
    def step(self, items, score="last", lm="none", resid_in=None, resid_out=False):
        """One decoder pass over several streams.

        items: list of dicts {storage: KVStorage, past: int (view length to append at), and either
                              embeds: bf16 [M,H] tensor  or  ids: list[int] (+ optional frames: bf16 [n,H] appended after)}
        score: 'last' (last row of each item), 'all', 'frame_ends' (item key 'score_rows': row offsets inside the item) or 'none'
        lm:    'none' or 'last' (lm_head logits for the last row of each item)
        Returns dict(head_logits [n,4], scores [n,2], lm_logits [n_lm,V] or None, views [CacheView per item]).

        Layer-pipeline stages (layer_range): a later stage passes resid_in = the fp32 residual stream [M,H] of the previous
        stage, with items that carry 'n_rows' instead of ids / frames; every stage but the last passes resid_out=True and
        gets dict(resid [M,H] fp32, views) back (score / lm still name the rows that are read: they select the precise rows)."""
        if (resid_in is None) != self.first_stage or bool(resid_out) == self.last_stage:
            raise _lib.MmdError(f"stage {self.layer_range}: resid_in is for later stages, resid_out for all but the last")
        with self._lock:
            return self._step_locked(items, score, lm, resid_in, resid_out)

    def _plan(self, items, score, lm):
        """Phase 1 of a step: validates every item and works out rows, positions and page needs WITHOUT touching any stream
        state.  Everything that can be refused (stale view, bad ids, bad shapes, context limit, KV pool exhausted) is refused
        here, before a rollback or a page allocation has happened."""
        H = self.cfg.hidden
        plan, emb_rows, seen = [], 0, set()
        pages_needed = 0
        for it in items:
            st, past = it["storage"], int(it["past"])
            if id(st) in seen:
                raise _lib.MmdError("the same stream appears twice in one step")
            seen.add(id(st))
            if past < 0 or past > st.length:
                raise _lib.MmdError("cache view is longer than the stream (stale view after a rollback)")
            chunk = None
            if it.get("n_rows") is not None:               # a later pipeline stage: the rows arrive as a residual stream
                rows = [0] * int(it["n_rows"])
            elif it.get("embeds") is not None:
                chunk = it["embeds"]
                rows = []
            else:
                ids = it.get("ids")
                rows = [int(i) for i in (ids if ids is not None else [])]
                if any(r < 0 or r >= self.cfg.vocab for r in rows):
                    raise _lib.MmdError("token id out of range")
                fr = it.get("frames")
                if fr is not None and fr.shape[0] > 0:
                    chunk = fr
            if chunk is not None:
                if chunk.dim() != 2 or chunk.shape[1] != H:
                    raise _lib.MmdError(f"embeddings must be [n, {H}], got {tuple(chunk.shape)}")
                rows = rows + [-(emb_rows + j) - 1 for j in range(chunk.shape[0])]
                emb_rows += chunk.shape[0]
            n_q = len(rows)
            if n_q == 0:
                raise _lib.MmdError("empty step item")
            new_len = past + n_q
            if new_len > self.max_context:
                raise _lib.MmdError(f"context {new_len} exceeds max_context {self.max_context}")
            if score == "frame_ends" and any(r < 0 or r >= n_q for r in it["score_rows"]):
                raise _lib.MmdError("score_rows outside the item")
            kept = min(len(st.pages), (past + PAGE - 1) // PAGE)          # pages that survive the rollback to `past`
            pages_needed += (new_len + PAGE - 1) // PAGE - kept - (len(st.pages) - kept)   # new pages minus recycled ones
            plan.append((it, st, past, rows, chunk, n_q, new_len))
        if pages_needed > len(self._free):
            import gc
            gc.collect()       # streams are released by KVStorage.__del__; cycles only die at a collection
            if pages_needed > len(self._free):
                raise _lib.MmdError(f"KV pool exhausted ({self.n_pages} pages of {PAGE} tokens, {len(self._free)} free, "
                                    f"{pages_needed} needed)")
        return plan

    def _meta_upload(self, meta):
        """One pinned staging buffer per in-flight step (ring of 8, grown on demand) instead of a cudaHostAlloc per step."""
        n = int(meta.size)
        i = self._meta_next
        self._meta_next = (i + 1) % len(self._meta_ring)
        slot = self._meta_ring[i]
        if slot is None or slot[0].numel() < n:
            cap = max(4096, 1 << (n - 1).bit_length())
            slot = [torch.empty(cap, dtype=torch.int32).pin_memory(), torch.empty(cap, dtype=torch.int32, device=self.device),
                    torch.cuda.Event()]
            self._meta_ring[i] = slot
        else:
            slot[2].synchronize()     # the copy issued 8 steps ago has long finished; this does not block in practice
        host, devbuf, ev = slot
        host.numpy()[:n] = meta
        devbuf[:n].copy_(host[:n], non_blocking=True)
        ev.record()
        return devbuf

    def _step_locked(self, items, score, lm, resid_in=None, resid_out=False):
        H, dev = self.cfg.hidden, self.device
        plan = self._plan(items, score, lm)
        src_row, tok_pos, tok_slot, desc, tables, score_rows, lm_rows = [], [], [], [], [], [], []
        emb_chunks = []
        q_start, max_n_q, max_kv = 0, 0, 0
        # phase 2: nothing below can be refused for a reason phase 1 could have seen.  Stream lengths are committed only
        # after the launch sequence was accepted; if it is not, every touched stream is left at its rollback target
        # (`past`): the append did not happen, and rows beyond `past` may have been overwritten, so they are not exposed.
        try:
            for it, st, past, rows, chunk, n_q, new_len in plan:
                if past < st.length:
                    st.truncate(past)  # an older view was passed back: rollback (all rollbacks first: they free pages)
            for it, st, past, rows, chunk, n_q, new_len in plan:
                st.ensure(new_len)
                if chunk is not None:
                    emb_chunks.append(chunk if chunk.dtype == torch.bfloat16 else chunk.to(torch.bfloat16))
                src_row += rows
                pg = st.pages
                for pos in range(past, new_len):
                    tok_pos.append(pos)
                    tok_slot.append(pg[pos // PAGE] * PAGE + pos % PAGE)
                desc += [q_start, n_q, new_len, len(tables)]
                tables += pg[:(new_len + PAGE - 1) // PAGE]
                if score == "last":
                    score_rows.append(q_start + n_q - 1)
                elif score == "all":
                    score_rows += list(range(q_start, q_start + n_q))
                elif score == "frame_ends":
                    score_rows += [q_start + r for r in it["score_rows"]]
                if lm == "last":
                    lm_rows.append(q_start + n_q - 1)
                elif lm == "all":
                    lm_rows += list(range(q_start, q_start + n_q))
                q_start += n_q
                max_n_q, max_kv = max(max_n_q, n_q), max(max_kv, new_len)
            M = q_start
            self._ensure_ws(M, len(lm_rows))
            # "precise rows" (csrc/api.cu): passes above 128 tokens re-run gate/up + down on bf16 hi+lo operands for the rows
            # whose outputs are read; shorter passes carry every row as hi+lo inside the main kernels
            prec = sorted(set(score_rows) | set(lm_rows)) if M > 128 else []
            if not (1 <= len(prec) <= 128):
                prec = []
            prec_of_row = []
            if prec:
                prec_of_row = [-1] * M
                for j, r in enumerate(prec):
                    prec_of_row[r] = j
            meta = np.asarray(src_row + tok_pos + tok_slot + desc + tables + score_rows + lm_rows + prec + prec_of_row, dtype=np.int32)
            meta_d = self._meta_upload(meta)
            base = meta_d.data_ptr()
            o = 0

            def seg(n):
                nonlocal o
                p = base + 4 * o
                o += n
                return p
            p_src, p_pos, p_slot = seg(M), seg(M), seg(M)
            p_desc, p_tab = seg(len(desc)), seg(len(tables))
            p_score, p_lm = seg(len(score_rows)), seg(len(lm_rows))
            p_prec, p_prec_of = seg(len(prec)), seg(len(prec_of_row))
            if emb_chunks:
                frame_tokens = emb_chunks[0] if len(emb_chunks) == 1 else torch.cat(emb_chunks, 0)
                frame_tokens = frame_tokens.contiguous()
            else:
                frame_tokens = None
            n_s, n_l = len(score_rows), len(lm_rows)
            head_logits = torch.empty(max(n_s, 1), 4, dtype=torch.float32, device=dev)
            scores = torch.empty(max(n_s, 1), 2, dtype=torch.float32, device=dev)
            lm_logits = torch.empty(n_l, self.cfg.vocab, dtype=torch.float32, device=dev) if n_l and not resid_out else None
            if resid_in is not None and (resid_in.shape != (M, H) or resid_in.dtype != torch.float32 or not resid_in.is_contiguous()
                                         or resid_in.device != dev):
                raise _lib.MmdError(f"resid_in must be a contiguous fp32 [{M}, {H}] tensor on {dev}")
            resid = torch.empty(M, H, dtype=torch.float32, device=dev) if resid_out else None
            if resid_out:
                n_l = 0
            step = _lib.Step(n_tokens=M, src_row=p_src, frame_tokens=_lib.ptr(frame_tokens), tok_pos=p_pos, tok_slot=p_slot,
                             n_streams=len(items), stream_desc=p_desc, block_tables=p_tab, max_n_q=max_n_q, max_kv_len=max_kv,
                             n_score_rows=n_s, score_rows=p_score, head_logits_out=head_logits.data_ptr(), scores_out=scores.data_ptr(),
                             n_lm_rows=n_l, lm_rows=p_lm, lm_logits_out=_lib.ptr(lm_logits),
                             n_prec_rows=len(prec), prec_rows=p_prec if prec else 0, prec_of_row=p_prec_of if prec else 0,
                             resid_in=_lib.ptr(resid_in), resid_out=_lib.ptr(resid))
            rc = self.lib.mmd_decoder_step(self.ctx, ctypes.byref(self.w), ctypes.byref(self.kv), ctypes.byref(step), self._ws.data_ptr(),
                                           self._ws.numel(), _lib.stream_ptr())
            _lib.check(rc, "mmd_decoder_step")
        except BaseException:
            for it, st, past, rows, chunk, n_q, new_len in plan:
                st.truncate(min(st.length, past))
            raise
        views = []
        for it, st, past, rows, chunk, n_q, new_len in plan:
            st.length = new_len
            views.append(CacheView(st, new_len))
        self._last_meta = (meta_d, frame_tokens, resid_in)  # keep alive until the kernels have consumed them
        if resid_out:
            return {"resid": resid, "views": views}
        return {"head_logits": head_logits[:n_s], "scores": scores[:n_s], "lm_logits": lm_logits, "views": views}
