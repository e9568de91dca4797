"""Thin Python wrappers over the C ABI (tensor -> pointer marshalling only; no arithmetic happens here)."""
import torch

from . import _lib
from ._lib import (ACT_GELU_ERF, ACT_GELU_TANH, ACT_NONE, EPI_BF16, EPI_F32, EPI_RESID_F32, EPI_SWIGLU_PAIR, EPI_T_F32,
                   EPI_T_SWIGLU, EPI_T_SWIGLU_IL)


def _chk2d(t, dtype):
    assert t.is_cuda and t.dtype == dtype and t.dim() == 2 and t.stride(1) == 1, (t.shape, t.dtype, t.stride())


def gemm(x, w, *, bias=None, act=ACT_NONE, out=None, epi=EPI_BF16):
    """Normal orientation: out[M,N] = epi(x[M,K] @ w[N,K]^T + bias)."""
    _chk2d(x, torch.bfloat16)
    _chk2d(w, torch.bfloat16)
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        assert epi in (EPI_BF16, EPI_F32)
        out = torch.empty(M, N, device=x.device, dtype=torch.bfloat16 if epi == EPI_BF16 else torch.float32)
    lib = _lib.load()
    rc = lib.mmd_gemm_bf16(_lib.context(x.device.index), epi, act, x.data_ptr(), 0, M, x.stride(0), w.data_ptr(), N,
                           w.stride(0), K, _lib.ptr(bias), out.data_ptr(), out.stride(0), 1, 0, _lib.stream_ptr())
    _lib.check(rc, "mmd_gemm_bf16")
    return out


def pack_blocked(w):
    """[N,K] row-major weights -> tile-blocked [N/128][K/64][128][64] (every 128x64 operand tile contiguous, 16 KB)."""
    N, K = w.shape
    assert N % 128 == 0 and K % 64 == 0, (N, K)
    return w.view(N // 128, 128, K // 64, 64).permute(0, 2, 1, 3).contiguous()


def gemm_t_partials(x, w, k_splits, out=None, blocked_shape=None):
    """Swap-AB split-K: returns fp32 partial planes [splits, M, N] of x[M,K] @ w[N,K]^T.
    blocked_shape=(N,K): `w` is a pack_blocked() buffer."""
    _chk2d(x, torch.bfloat16)
    M, K = x.shape
    if blocked_shape is not None:
        N = blocked_shape[0]
        assert blocked_shape[1] == K
        ldw = -1
    else:
        _chk2d(w, torch.bfloat16)
        N = w.shape[0]
        ldw = w.stride(0)
    lib = _lib.load()
    splits = lib.mmd_gemm_splits(K, k_splits)
    if out is None:
        out = torch.empty(splits, M, N, device=x.device, dtype=torch.float32)
    rc = lib.mmd_gemm_bf16(_lib.context(x.device.index), EPI_T_F32, ACT_NONE, w.data_ptr(), 0, N, ldw,
                           x.data_ptr(), M, x.stride(0), K, 0, out.data_ptr(), out.stride(1), k_splits, out.stride(0),
                           _lib.stream_ptr())
    _lib.check(rc, "mmd_gemm_bf16(T_F32)")
    return out


def gemm_t_swiglu_interleaved(x, w_il, out=None):
    """Swap-AB fused SwiGLU on ONE interleaved weight matrix (row 2j = gate_j, row 2j+1 = up_j): out[M, N] with N = rows/2."""
    _chk2d(x, torch.bfloat16)
    _chk2d(w_il, torch.bfloat16)
    M, K = x.shape
    N = w_il.shape[0] // 2
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    lib = _lib.load()
    rc = lib.mmd_gemm_bf16(_lib.context(x.device.index), EPI_T_SWIGLU_IL, ACT_NONE, w_il.data_ptr(), 0, 2 * N, w_il.stride(0),
                           x.data_ptr(), M, x.stride(0), K, 0, out.data_ptr(), out.stride(0), 1, 0, _lib.stream_ptr())
    _lib.check(rc, "mmd_gemm_bf16(T_SWIGLU_IL)")
    return out


def gemm_t_swiglu(x, w_gate, w_up, out=None, blocked_shape=None):
    """Swap-AB fused SwiGLU: out[M,N] = silu(x @ w_gate^T) * (x @ w_up^T), bf16."""
    _chk2d(x, torch.bfloat16)
    M, K = x.shape
    if blocked_shape is not None:
        N, ldw = blocked_shape[0], -1
    else:
        _chk2d(w_gate, torch.bfloat16)
        _chk2d(w_up, torch.bfloat16)
        N, ldw = w_gate.shape[0], w_gate.stride(0)
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    lib = _lib.load()
    rc = lib.mmd_gemm_bf16(_lib.context(x.device.index), EPI_T_SWIGLU, ACT_NONE, w_gate.data_ptr(), w_up.data_ptr(), N,
                           ldw, x.data_ptr(), M, x.stride(0), K, 0, out.data_ptr(), out.stride(0), 1, 0,
                           _lib.stream_ptr())
    _lib.check(rc, "mmd_gemm_bf16(T_SWIGLU)")
    return out


def gemm_f32_planes(x, w, k_splits=1, out=None):
    """Normal orientation, fp32 split-K partial planes [splits, M, N] of x[M,K] @ w[N,K]^T.  With M >= 1024 this is the
    CTA-pair kernel writing plane `ks` through a 3-D TMA store: the decoder's q/k/v, o (1 plane) and down (3 planes)
    projections of passes with >= 1024 tokens (csrc/api.cu, mmd_decoder_step)."""
    _chk2d(x, torch.bfloat16)
    _chk2d(w, torch.bfloat16)
    M, K = x.shape
    N = w.shape[0]
    lib = _lib.load()
    splits = lib.mmd_gemm_splits(K, k_splits)
    if out is None:
        out = torch.empty(splits, M, N, device=x.device, dtype=torch.float32)
    rc = lib.mmd_gemm_bf16(_lib.context(x.device.index), EPI_F32, ACT_NONE, x.data_ptr(), 0, M, x.stride(0), w.data_ptr(), N,
                           w.stride(0), K, 0, out.data_ptr(), out.stride(1), k_splits, out.stride(0), _lib.stream_ptr())
    _lib.check(rc, "mmd_gemm_bf16(F32 planes)")
    return out


def gemm_swiglu_pair(x, w_il, out=None):
    """Normal orientation (CTA-pair kernel) on the interleaved gate/up matrix (row 2j = gate_j, 2j+1 = up_j):
    out[M, N] = silu(x @ gate^T) * (x @ up^T) with N = rows/2 — the decoder's gate/up of passes with >= 2048 tokens."""
    _chk2d(x, torch.bfloat16)
    _chk2d(w_il, torch.bfloat16)
    M, K = x.shape
    N = w_il.shape[0] // 2
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    lib = _lib.load()
    rc = lib.mmd_gemm_bf16(_lib.context(x.device.index), EPI_SWIGLU_PAIR, ACT_NONE, x.data_ptr(), 0, M, x.stride(0), w_il.data_ptr(),
                           2 * N, w_il.stride(0), K, 0, out.data_ptr(), out.stride(0), 1, 0, _lib.stream_ptr())
    _lib.check(rc, "mmd_gemm_bf16(SWIGLU_PAIR)")
    return out


def split_hilo(x):
    """fp32 [M, K] -> bf16 [M, 2K] = [hi | lo] with hi = bf16(x), lo = bf16(x - hi) (plumbing for the tests; on the path the
    RMSNorm / SwiGLU epilogues write this layout themselves)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], 1).contiguous()


def gemm_t_partials_hilo(x2, w, k_splits, out=None):
    """Swap-AB split-K on a hi+lo activation pair x2 [M <= 128, 2K]: fp32 planes [splits, M, N] of (hi + lo) @ w[N,K]^T."""
    _chk2d(x2, torch.bfloat16)
    _chk2d(w, torch.bfloat16)
    M, K = x2.shape[0], w.shape[1]
    assert x2.shape[1] == 2 * K
    N = w.shape[0]
    lib = _lib.load()
    splits = lib.mmd_gemm_splits(K, k_splits)
    if out is None:
        out = torch.empty(splits, M, N, device=x2.device, dtype=torch.float32)
    rc = lib.mmd_gemm_bf16(_lib.context(x2.device.index), EPI_T_F32 | _lib.GEMM_Y_HILO, ACT_NONE, w.data_ptr(), 0, N, w.stride(0),
                           x2.data_ptr(), M, x2.stride(0), K, 0, out.data_ptr(), out.stride(1), k_splits, out.stride(0), _lib.stream_ptr())
    _lib.check(rc, "mmd_gemm_bf16(T_F32 | Y_HILO)")
    return out


def gemm_t_swiglu_hilo(x2, w_il, out=None):
    """Swap-AB fused SwiGLU on a hi+lo activation pair x2 [M <= 128, 2K] against the interleaved gate/up matrix w_il [2N, K]
    (row 2j = gate_j, 2j+1 = up_j); the output is a hi+lo pair too: [M, 2N] = [bf16(v) | bf16(v - bf16(v))]."""
    _chk2d(x2, torch.bfloat16)
    _chk2d(w_il, torch.bfloat16)
    M, K = x2.shape[0], w_il.shape[1]
    assert x2.shape[1] == 2 * K
    N = w_il.shape[0] // 2
    if out is None:
        out = torch.empty(M, 2 * N, device=x2.device, dtype=torch.bfloat16)
    lib = _lib.load()
    up = w_il[1:]     # up rows start one row further; both operands use row stride 2K
    rc = lib.mmd_gemm_bf16(_lib.context(x2.device.index), EPI_T_SWIGLU | _lib.GEMM_Y_HILO | _lib.GEMM_OUT_HILO, ACT_NONE, w_il.data_ptr(),
                           up.data_ptr(), N, 2 * w_il.stride(0), x2.data_ptr(), M, x2.stride(0), K, 0, out.data_ptr(), out.stride(0), 1, 0,
                           _lib.stream_ptr())
    _lib.check(rc, "mmd_gemm_bf16(T_SWIGLU | Y_HILO | OUT_HILO)")
    return out
