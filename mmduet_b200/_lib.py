"""ctypes binding of libmmduet_b200.so.  There is deliberately no fallback: a missing library or a non-B200 device
raises, so a silent PyTorch path can never stand in for the CUDA kernels."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmduet_b200.so")

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

EPI_BF16, EPI_RESID_F32, EPI_T_F32, EPI_T_SWIGLU, EPI_F32 = 0, 1, 2, 3, 4
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF = 0, 1, 2


class MmdError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()
_ctx = {}

# name -> (restype, argtypes); every symbol include/mmduet_b200.h declares must be listed here
SIGNATURES = {
    "mmd_version": (ctypes.c_char_p, []),
    "mmd_last_error": (ctypes.c_char_p, []),
    "mmd_create": (c_void_p, [c_int]),
    "mmd_destroy": (None, [c_void_p]),
    "mmd_gemm_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                              c_int64, c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p]),
    "mmd_gemm_splits": (c_int, [c_int64, c_int]),
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


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()
