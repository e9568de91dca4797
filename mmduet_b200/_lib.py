"""ctypes binding of libmmduet_b200.so.  There is deliberately no fallback: a missing library or a non-B200 device
raises, so a silent PyTorch path can never stand in for the CUDA kernels."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMD_LIB_PATH") or os.path.join(_HERE, "libmmduet_b200.so")   # override: debug builds only

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

EPI_BF16, EPI_RESID_F32, EPI_T_F32, EPI_T_SWIGLU, EPI_F32, EPI_BF16_HILO = 0, 1, 2, 3, 4, 5
EPI_SWIGLU_PAIR, EPI_T_SWIGLU_IL = 6, 7
GEMM_Y_HILO, GEMM_OUT_HILO = 0x100, 0x200
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF = 0, 1, 2


class MmdError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()
_ctx = {}

PAGE_TOKENS = 64
DT_U8, DT_BF16, DT_F32 = 0, 1, 2
c_float_p = ctypes.c_void_p  # device pointers travel as integers


class VitLayer(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_w", "ln2_b",
                                        "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class VitWeights(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("image_size", "patch_size", "dim", "heads", "mlp", "n_layers", "k_pad", "attn_out_split")] + \
               [("patch_w", c_void_p), ("patch_b", c_void_p), ("pos_emb", c_void_p), ("layers", ctypes.POINTER(VitLayer))]


class ProjectorWeights(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("vit_dim", "hidden", "n_src_tokens", "n_gather", "n_out", "max_taps", "maxpool", "hilo")] + \
               [(n, c_void_p) for n in ("w1", "b1", "w2", "b2", "gather_idx", "tap_idx", "tap_w")] + \
               [("pool_group", c_int), ("pool_gather_idx", c_void_p), ("pool_row_w", c_void_p)]


class DecLayer(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1_w", "qkv_w", "qkv_b", "o_w", "ln2_w", "gate_up_w", "down_w")]


class DecWeights(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("hidden", "n_layers", "q_heads", "kv_heads", "head_dim", "mlp", "vocab", "max_pos")] + \
               [("rms_eps", c_float), ("layers", ctypes.POINTER(DecLayer))] + \
               [(n, c_void_p) for n in ("final_norm_w", "embed", "lm_head", "heads_w", "rope_cos", "rope_sin")]


class KvPool(ctypes.Structure):
    _fields_ = [("pool", c_void_p), ("layer_stride", c_int64), ("n_pages", c_int)]


class Step(ctypes.Structure):
    _fields_ = [("n_tokens", c_int), ("src_row", c_void_p), ("frame_tokens", c_void_p), ("tok_pos", c_void_p),
                ("tok_slot", c_void_p), ("n_streams", c_int), ("stream_desc", c_void_p), ("block_tables", c_void_p),
                ("max_n_q", c_int), ("max_kv_len", c_int), ("n_score_rows", c_int), ("score_rows", c_void_p),
                ("head_logits_out", c_void_p), ("scores_out", c_void_p), ("n_lm_rows", c_int), ("lm_rows", c_void_p),
                ("lm_logits_out", c_void_p), ("n_prec_rows", c_int), ("prec_rows", c_void_p), ("prec_of_row", c_void_p),
                ("resid_in", c_void_p), ("resid_out", c_void_p)]


P = ctypes.POINTER

# name -> (restype, argtypes); every symbol include/mmduet_b200.h declares must be listed here
SIGNATURES = {
    "mmd_version": (ctypes.c_char_p, []),
    "mmd_last_error": (ctypes.c_char_p, []),
    "mmd_create": (c_void_p, [c_int]),
    "mmd_destroy": (None, [c_void_p]),
    "mmd_gemm_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                              c_int64, c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p]),
    "mmd_gemm_splits": (c_int, [c_int64, c_int]),
    "mmd_num_sms": (c_int, [c_void_p]),
    "mmd_set_attention_impl": (c_int, [c_int]),
    "mmd_set_gemm_2cta": (c_int, [c_int]),
    "mmd_launch_count": (ctypes.c_ulonglong, [c_void_p]),
    "mmd_profile_num_tags": (c_int, []),
    "mmd_profile_tag_name": (ctypes.c_char_p, [c_int]),
    "mmd_profile_start": (c_int, [c_void_p, ctypes.c_char_p]),
    "mmd_profile_stop": (c_int, [c_void_p, c_void_p, c_void_p]),
    "mmd_frame_ingest": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mmd_im2col": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mmd_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_float, c_void_p]),
    "mmd_vit_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mmd_resid_add_rmsnorm": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                      c_float, c_void_p]),
    "mmd_resid_add_rmsnorm_precise": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                              c_float, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "mmd_final_norm_heads": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "mmd_qkv_finish": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mmd_kv_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mmd_kv_attention_splits": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "mmd_tap_pool": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_void_p]),
    "mmd_heads": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mmd_probe_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mmd_argmax": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "mmd_grounding_sweep": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "mmd_vit_workspace_bytes": (c_int64, [P(VitWeights), c_int]),
    "mmd_vit_forward": (c_int, [c_void_p, P(VitWeights), c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int64,
                                c_void_p]),
    "mmd_projector_workspace_bytes": (c_int64, [P(ProjectorWeights), c_int]),
    "mmd_projector_pool": (c_int, [c_void_p, P(ProjectorWeights), c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64,
                                   c_void_p]),
    "mmd_decoder_workspace_bytes": (c_int64, [c_void_p, P(DecWeights), c_int, c_int]),
    "mmd_decoder_step": (c_int, [c_void_p, P(DecWeights), P(KvPool), P(Step), c_void_p, c_int64, c_void_p]),
}


def load():
    """Loads the shared library (building nothing: run mmduet_b200/build.py or __graft_entry__.build() first)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise MmdError(f"{LIB_PATH} is missing: build it with `python -m mmduet_b200.build` "
                               "(nvcc, sm_100a). There is no CPU/PyTorch fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def last_error():
    return load().mmd_last_error().decode()


def check(rc, what=""):
    if rc != 0:
        raise MmdError(f"{what} failed ({rc}): {last_error()}")


def context(device=None):
    """Per-device mmd_ctx handle (created lazily; requires a B200)."""
    import torch
    if device is None:
        device = torch.cuda.current_device()
    lib = load()
    with _lock:
        h = _ctx.get(device)
        if h is None:
            h = lib.mmd_create(int(device))
            if not h:
                raise MmdError(f"mmd_create({device}) failed: {lib.mmd_last_error().decode()}")
            _ctx[device] = h
    return h


def launch_count(device=None):
    return int(load().mmd_launch_count(context(device)))


def profile_start(tags="all", device=None):
    check(load().mmd_profile_start(context(device), tags.encode()), "mmd_profile_start")


def profile_stop(device=None):
    """Returns {tag: (total_ms, launches)} for the tags that fired."""
    lib = load()
    n = lib.mmd_profile_num_tags()
    ms = (ctypes.c_float * n)()
    cnt = (ctypes.c_int * n)()
    check(lib.mmd_profile_stop(context(device), ms, cnt), "mmd_profile_stop")
    return {lib.mmd_profile_tag_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i] > 0}


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()
