"""TEST INFRASTRUCTURE ONLY — imports the reference's OWN model class from /root/reference through three shims.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
This module works only where /root/reference exists (the authoring container); it is used to (1) validate the
restatement in oracle/restate.py and (2) generate the golden vectors under tests/golden/ (oracle/make_golden.py).

Shims (none of them touches the arithmetic of the reference files):
  1. `peft`  — absent here; models/modeling_live.py:3 imports LoraConfig/get_peft_model/PeftModel at module import.
  2. `llava.model.llava_arch.LlavaMetaModel` — LLaVA-NeXT is an unpinned git dependency (README.md:31-36) that is not
     vendored.  The stub restates what its published code does for `llava-onevision-qwen2-7b-ov`:
       * SigLipVisionTower: SigLIP-so400m/14@384 vision model with the LAST encoder layer deleted and head=Identity;
         forward returns hidden_states[-1], i.e. the output of the 26th layer BEFORE post_layernorm
         [recalled from llava/model/multimodal_encoder/siglip_encoder.py, unpinned];
       * mm_projector 'mlp2x_gelu': Linear(1152,3584) -> nn.GELU() -> Linear(3584,3584)
         [recalled from llava/model/multimodal_projector/builder.py, unpinned].
  3. transformers 5.x guard: video_head_live_llava_qwen.py:73 sets `config.rope_scaling = None`, which on the
     installed transformers 5.5 wipes `rope_parameters`; the guard ignores that one assignment.
"""
import importlib.machinery
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("MMDUET_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


class _SigLipTowerStub(nn.Module):
    """Restates LLaVA-NeXT's SigLipVisionTower.forward for a batched tensor input."""

    def __init__(self, vision_cfg):
        super().__init__()
        from transformers import SiglipVisionModel
        self.vision_tower = SiglipVisionModel(vision_cfg)
        del self.vision_tower.vision_model.encoder.layers[-1:]
        self.vision_tower.vision_model.head = nn.Identity()
        self.vision_tower.requires_grad_(False)
        self.num_patches_per_side = vision_cfg.image_size // vision_cfg.patch_size
        self.image_processor = None

    def forward(self, images):
        vm = self.vision_tower.vision_model
        x = vm.embeddings(images.to(dtype=vm.embeddings.patch_embedding.weight.dtype))
        out = vm.encoder(inputs_embeds=x)
        return out.last_hidden_state.to(images.dtype)  # pre-post_layernorm, all patch tokens


def _install_stubs():
    if "peft" not in sys.modules:
        peft = types.ModuleType("peft")
        peft.__spec__ = importlib.machinery.ModuleSpec("peft", None)

        class _Unavailable:
            def __init__(self, *a, **k):
                raise RuntimeError("peft is not installed; the oracle runs without LoRA")

            @classmethod
            def from_pretrained(cls, *a, **k):
                raise RuntimeError("peft is not installed; the oracle runs without LoRA")

        peft.LoraConfig = _Unavailable
        peft.PeftModel = _Unavailable
        peft.get_peft_model = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("peft is not installed"))
        sys.modules["peft"] = peft

    if "llava.model.llava_arch" not in sys.modules:
        llava = types.ModuleType("llava")
        llava.__path__ = []
        llava_model = types.ModuleType("llava.model")
        llava_model.__path__ = []
        arch = types.ModuleType("llava.model.llava_arch")

        class LlavaMetaModel:
            def __init__(self, config):
                super().__init__(config)
                vcfg = getattr(config, "oracle_vision_config", None)
                if vcfg is None:
                    raise RuntimeError("config.oracle_vision_config (SiglipVisionConfig) must be set for the stub")
                self.vision_tower = _SigLipTowerStub(vcfg)
                self.mm_projector = nn.Sequential(
                    nn.Linear(vcfg.hidden_size, config.hidden_size), nn.GELU(),
                    nn.Linear(config.hidden_size, config.hidden_size))

            def get_vision_tower(self):
                return self.vision_tower

        arch.LlavaMetaModel = LlavaMetaModel
        sys.modules["llava"] = llava
        sys.modules["llava.model"] = llava_model
        sys.modules["llava.model.llava_arch"] = arch

    if "models" not in sys.modules or not getattr(sys.modules["models"], "__oracle_shim__", False):
        pkg = types.ModuleType("models")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "models")]
        pkg.__oracle_shim__ = True
        sys.modules["models"] = pkg
        sub = types.ModuleType("models.live_llava")
        sub.__path__ = [os.path.join(REFERENCE_ROOT, "models", "live_llava")]
        sys.modules["models.live_llava"] = sub


def import_reference():
    """Returns the reference modules (vision_live, modeling_live, video_head_live_llava_qwen)."""
    if not reference_available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (it never is on the GPU box)")
    _install_stubs()
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vl = importlib.import_module("models.vision_live")
        ml = importlib.import_module("models.modeling_live")
        vh = importlib.import_module("models.live_llava.video_head_live_llava_qwen")
    cfg_cls = vh.VideoHeadLiveLlavaQwenConfig
    if not getattr(cfg_cls, "__oracle_rope_guard__", False):
        orig_setattr = cfg_cls.__setattr__

        def guarded(self, key, value):
            if key == "rope_scaling" and value is None:
                return  # shim 3
            return orig_setattr(self, key, value)

        cfg_cls.__setattr__ = guarded
        cfg_cls.__oracle_rope_guard__ = True
    return vl, ml, vh


def build_reference_model(arch, state_dict=None, dtype=torch.float32, seed=1234):
    """Instantiates the reference's VideoHeadLiveLlavaQwenForCausalLM (random init or from `state_dict`).

    `arch` is an oracle.arch.Arch.  Returns the nn.Module in eval mode on CPU."""
    from transformers import SiglipVisionConfig
    _, _, vh = import_reference()
    vcfg = SiglipVisionConfig(hidden_size=arch.vit_dim, intermediate_size=arch.vit_mlp,
                              num_hidden_layers=arch.vit_layers_total, num_attention_heads=arch.vit_heads,
                              image_size=arch.image_size, patch_size=arch.patch_size,
                              hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)
    cfg = vh.VideoHeadLiveLlavaQwenConfig(
        hidden_size=arch.hidden, intermediate_size=arch.mlp, num_hidden_layers=arch.layers,
        num_attention_heads=arch.q_heads, num_key_value_heads=arch.kv_heads, vocab_size=arch.vocab,
        rms_norm_eps=arch.rms_eps, rope_theta=arch.rope_theta, max_position_embeddings=arch.max_pos,
        tie_word_embeddings=False, attn_implementation="sdpa",
        video_pooling_stride=arch.pool_stride, frame_num_tokens=arch.frame_tokens,
        frame_resolution=arch.image_size)
    cfg.mm_spatial_pool_mode = arch.pool_mode
    # fields VideoHeadLiveConfigMixin.__init__ sets from the CLI flags (configuration_live.py:21-35); the installed
    # transformers' Qwen2Config does not chain to the mixin's __init__, so they are set here
    for k, v in dict(frame_num_tokens=arch.frame_tokens, frame_resolution=arch.image_size, v_placeholder="<image>",
                     v_placeholder_id=None, frame_token_cls=False, frame_token_pooled=[7, 7]).items():
        if getattr(cfg, k, None) is None:
            setattr(cfg, k, v)
    cfg.oracle_vision_config = vcfg
    torch.manual_seed(seed)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = vh.VideoHeadLiveLlavaQwenForCausalLM(cfg)
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        # `vision_encoder.*` aliases `model.vision_tower.*` (video_head_live_llava_qwen.py:82: same module, second name)
        missing = [m for m in missing if "post_layernorm" not in m and "position_ids" not in m
                   and not m.startswith("vision_encoder.")]
        if missing or unexpected:
            raise RuntimeError(f"state_dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
    model = model.to(dtype).eval()
    model.requires_grad_(False)
    return model


# ------------------------------------------------------------------------------------------------------------
# The reference's OWN frame-loop classes (test/inference.py:20-313 LiveInferForBenchmark, demo/liveinfer.py:60-105
# LiveInferForDemo) on CPU.  Extra shims, none of which touches the loop's logic:
#   4. `llava.mm_utils / llava.model.builder / llava.constants / llava.conversation` (test/inference.py:11-14) and
#      `torchvision.io.read_video` (:7, gone from the installed torchvision): imported at module level, never used by the
#      live loop -> empty name stubs.
#   5. `models.build_model_and_tokenizer` (test/inference.py:16,24) returns the model from build_reference_model() and the
#      tokenizer handed to build_reference_loop(), instead of downloading lmms-lab/llava-onevision-qwen2-7b-ov.
#   6. device: the loop hard-codes 'cuda' (torch.zeros(..., device='cuda'), .to('cuda')).  The module's `torch` global is
#      replaced by a forwarding proxy that maps device='cuda' to 'cpu' for torch.zeros / torch.tensor, and tensors that
#      the loop moves with .to('cuda') (tokenizer output, pixel values) are a Tensor subclass whose .to() does the same.
#   7. cache semantics of the PINNED transformers==4.44.2 (requirements.txt:41): with past_key_values=None on the first
#      call, Qwen2Model returns LEGACY TUPLE caches (immutable: every forward builds a new tuple), which is what makes
#      `remove_assistant_turns` a rollback when the cache returned by fast_greedy_generate is dropped
#      (test/inference.py:265-269).  The installed transformers 5.5 mutates a DynamicCache in place, so
#      the model handed to the loop copies the cache it is given before each forward (LegacyCacheModel).
#   8. LLaVA's SigLipImageProcessor (get_vision_tower().image_processor, test/inference.py:27,203) restated for frames
#      that are already 384x384: rescale 1/255, normalise mean = std = 0.5 [recalled from
#      llava/model/multimodal_encoder/siglip_encoder.py, unpinned].
# ------------------------------------------------------------------------------------------------------------
class CpuTensor(torch.Tensor):
    """A tensor whose .to('cuda') / .cuda() stay on the CPU (shim 6)."""

    def to(self, *args, **kwargs):
        args = tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) else a for a in args)
        if isinstance(kwargs.get("device"), str) and kwargs["device"].startswith("cuda"):
            kwargs["device"] = "cpu"
        return super().to(*args, **kwargs)

    def cuda(self, *a, **k):
        return self


def as_cpu_tensor(t):
    return t.as_subclass(CpuTensor)


class _TorchProxy:
    """Forwards everything to torch; zeros / tensor with device='cuda' are created on the CPU (shim 6)."""

    def __init__(self):
        self.__dict__["_t"] = torch

    def __getattr__(self, name):
        return getattr(self._t, name)

    @staticmethod
    def _dev(kwargs):
        if isinstance(kwargs.get("device"), str) and kwargs["device"].startswith("cuda"):
            kwargs["device"] = "cpu"
        return kwargs

    def zeros(self, *a, **k):
        return torch.zeros(*a, **self._dev(k))

    def tensor(self, *a, **k):
        return torch.tensor(*a, **self._dev(k))


class SigLipImageProcessorStub:
    """Shim 8."""
    image_mean, image_std, rescale_factor, size = (0.5, 0.5, 0.5), (0.5, 0.5, 0.5), 0.00392156862745098, (384, 384)

    def preprocess(self, images, return_tensors="pt"):
        x = torch.as_tensor(images)
        assert tuple(x.shape[-2:]) == self.size, "the stub only handles frames already padded to 384x384 (test/datasets.py)"
        x = (x.float() * self.rescale_factor - 0.5) / 0.5
        return {"pixel_values": as_cpu_tensor(x)}


class TokenizerOnCpu:
    """Wraps any tokenizer: apply_chat_template(..., return_tensors='pt') returns the ids as a [1, n] CpuTensor, the
    transformers==4.44.2 return type (5.x returns a BatchEncoding unless return_dict=False)."""

    def __init__(self, tok):
        self._tok = tok
        self.decoded = []          # raw id lists handed to decode(): the generated responses, in order

    def __getattr__(self, name):
        return getattr(self._tok, name)

    def decode(self, ids, **kw):
        self.decoded.append([int(i) for i in (ids.tolist() if torch.is_tensor(ids) else ids)])
        return self._tok.decode(ids, **kw)

    def apply_chat_template(self, conversation, **kw):
        try:
            out = self._tok.apply_chat_template(conversation, return_dict=False, **kw)
        except TypeError:
            out = self._tok.apply_chat_template(conversation, **kw)
        if not torch.is_tensor(out):
            out = torch.as_tensor(out["input_ids"] if hasattr(out, "keys") else out, dtype=torch.long)
        return as_cpu_tensor(out.view(1, -1))


class LegacyCacheModel:
    """Shim 7: hands the reference loop a model whose returned caches are never mutated by later calls."""

    def __init__(self, model):
        self.__dict__["_m"] = model
        self.__dict__["calls"] = []    # (tokens in the call, top-2 gap of the last position's lm logits)

    def __getattr__(self, name):
        return getattr(self._m, name)

    def eval(self):
        self._m.eval()
        return self

    def __call__(self, *args, past_key_values=None, **kwargs):
        import copy
        if past_key_values is not None:
            past_key_values = copy.deepcopy(past_key_values)
        # plain tensors inside the model (the device shim's Tensor subclass must not end up in the cache)
        kwargs = {k: (v.as_subclass(torch.Tensor) if torch.is_tensor(v) else v) for k, v in kwargs.items()}
        out = self._m(*args, past_key_values=past_key_values, **kwargs)
        t = out.logits[0, -1].float().topk(2).values
        self.calls.append((int(kwargs["inputs_embeds"].shape[1]), float(t[0] - t[1])))
        return out


def import_reference_loop():
    """Returns (test.inference, demo.liveinfer) of the reference, importable on CPU."""
    _, ml, vh = import_reference()
    for name in ("llava.mm_utils", "llava.model.builder", "llava.constants", "llava.conversation"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["llava.mm_utils"].tokenizer_image_token = None
    sys.modules["llava.model.builder"].load_pretrained_model = None
    sys.modules["llava.constants"].IMAGE_TOKEN_INDEX, sys.modules["llava.constants"].DEFAULT_IMAGE_TOKEN = -200, "<image>"
    sys.modules["llava.conversation"].conv_templates = {}
    import torchvision.io as tvio
    if not hasattr(tvio, "read_video"):      # removed from recent torchvision; only the unused load_video() calls it
        tvio.read_video = None
    models = sys.modules["models"]
    models.fast_greedy_generate = ml.fast_greedy_generate
    models.parse_args = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("CLI parsing is not part of the loop"))
    if not hasattr(models, "build_model_and_tokenizer"):
        models.build_model_and_tokenizer = lambda **k: (_ for _ in ()).throw(RuntimeError("use build_reference_loop()"))
    for pkg in ("test", "demo"):
        if pkg not in sys.modules or not getattr(sys.modules[pkg], "__oracle_shim__", False):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REFERENCE_ROOT, pkg)]
            m.__oracle_shim__ = True
            sys.modules[pkg] = m
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ti = importlib.import_module("test.inference")
        dl = importlib.import_module("demo.liveinfer")
    proxy = _TorchProxy()
    ti.torch = proxy
    dl.torch = proxy
    return ti, dl


def build_reference_loop(arch, state_dict, tokenizer, args, demo=False, dtype=torch.float32):
    """The reference's LiveInferForBenchmark / LiveInferForDemo (unmodified class bodies) over the reference's own model class
    with `state_dict`, on CPU.  `args`: any dataclass with the reference's flag names (mmduet_b200.arguments_live.
    LiveTestArguments has the same names and defaults); bf16/fp16 are forced off so the loop runs in `dtype`."""
    import contextlib
    import dataclasses
    import io
    ti, dl = import_reference_loop()
    model = build_reference_model(arch, state_dict, dtype=dtype)
    model.get_vision_tower().image_processor = SigLipImageProcessorStub()
    model.config.eos_token_id = getattr(tokenizer, "eos_token_id", None)
    wrapped = LegacyCacheModel(model)
    tok = TokenizerOnCpu(tokenizer)
    sys.modules["models"].build_model_and_tokenizer = lambda **k: (wrapped, tok)
    ti.build_model_and_tokenizer = sys.modules["models"].build_model_and_tokenizer
    args = dataclasses.replace(args, bf16=False, fp16=False)
    cls = dl.LiveInferForDemo if demo else ti.LiveInferForBenchmark
    with contextlib.redirect_stdout(io.StringIO()):
        loop = cls(args)
    assert loop.torch_dtype == torch.float32
    loop.torch_dtype = dtype
    return loop
