/* mmduet_b200 — C ABI of the B200-native per-frame streaming hot path of MMDuet.
 *
 * The reference (yellow-binary-tree/MMDuet) is pure Python and has no FFI; its boundary is the Python surface in
 * models/vision_live.py, models/modeling_live.py, models/live_llava/video_head_live_llava_qwen.py and
 * test/inference.py.  The functions below are what a ctypes binding underneath that surface calls; each one names
 * the reference call site whose arithmetic it replaces.  Conventions: every pointer is a caller-owned CUDA device
 * pointer unless stated otherwise, sizes/strides are int64_t in ELEMENTS, `stream` is a cudaStream_t passed as
 * void*, calls are asynchronous with respect to the host, and the return value is 0 on success or a negative
 * MMD_ERR_* code with a thread-local message available from mmd_last_error().  There is no CPU fallback.
 */
#ifndef MMDUET_B200_H_
#define MMDUET_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MMD_API __attribute__((visibility("default")))
#else
#define MMD_API
#endif

#define MMD_OK 0
#define MMD_ERR_ARG (-2)
#define MMD_ERR_CUDA (-5)

typedef struct mmd_ctx mmd_ctx;

MMD_API const char* mmd_version(void);
MMD_API const char* mmd_last_error(void);
/* One context per device/process; fails (NULL) on anything that is not sm_100. */
MMD_API mmd_ctx* mmd_create(int device);
MMD_API void mmd_destroy(mmd_ctx*);

/* epilogue / activation selectors of mmd_gemm_bf16 */
#define MMD_EPI_BF16 0
#define MMD_EPI_RESID_F32 1
#define MMD_EPI_T_F32 2
#define MMD_EPI_T_SWIGLU 3
#define MMD_EPI_F32 4
#define MMD_ACT_NONE 0
#define MMD_ACT_GELU_TANH 1
#define MMD_ACT_GELU_ERF 2

/* D[i, j] = sum_k X[i, k] * Y[j, k], bf16 operands (both K-major), fp32 accumulation in TMEM (tcgen05.mma).
 * Replaces every nn.Linear on the path (cuBLAS in the reference): SigLIP q/k/v/out/fc1/fc2 and the patch-embed
 * Conv2d as a GEMM (video_head_live_llava_qwen.py:96-98), mm_projector (:90-91) and the Qwen2 projections
 * (:141-150).  Normal epilogues: X = activations [M,K], Y = weights [N,K], out[M,N].  T epilogues (swap-AB for small
 * M): X (and X2) = weights [N,K], Y = activations [M,K], out[M,N]; MMD_EPI_T_F32 writes `k_splits` fp32 partial
 * planes `split_stride` elements apart (see mmd_gemm_splits for the effective plane count). */
MMD_API int mmd_gemm_bf16(mmd_ctx*, int epi, int act, const void* X, const void* X2, int64_t x_rows, int64_t ldx,
                  const void* Y, int64_t y_rows, int64_t ldy, int64_t K, const float* bias, void* out, int64_t ldo,
                  int k_splits, int64_t split_stride, void* stream);
MMD_API int mmd_gemm_splits(int64_t K, int k_splits);

#ifdef __cplusplus
}
#endif
#endif /* MMDUET_B200_H_ */
