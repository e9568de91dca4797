// Score post-processing of the reference's grounding evaluator on the device:
//   test/evaluate.py:166-167  smooth_pred_list   : mean over the window [i-w, i+w] clipped to the list (np.mean)
//   test/evaluate.py:170-173  normalize_pred_list: (p - min) / (max - min)
//   test/evaluate.py:129-137  calculate_iou      : pred = p >= threshold; |pred & gold| / |pred | gold|
//   test/evaluate.py:374-392  the sweep          : smoothing windows 0..14 x thresholds np.arange(0.30, 0.71, 0.02)
// Everything is float64 and follows numpy's summation order (pairwise_sum: < 8 elements sequential; otherwise eight
// running sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail), so the normalised scores are
// bit-identical to the evaluator's and the intersection / union counts are exact integers.
#include "../../include/mmduet_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ double numpy_pairwise_sum(const double* a, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r += a[i];
    return r;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] += a[i + j];
  }
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += a[i];
  return res;
}

// grid (videos, windows); one block per (video, smoothing window)
__global__ void grounding_sweep_kernel(const double* __restrict__ scores, const uint8_t* __restrict__ gold, const int* __restrict__ lens,
                                       int t_max, const int* __restrict__ windows, const double* __restrict__ thresholds, int n_thr,
                                       int* __restrict__ counts, double* __restrict__ norm_out, int* __restrict__ degenerate) {
  extern __shared__ double sm[];                 // t_max smoothed scores
  __shared__ double red_min[32], red_max[32];
  __shared__ int cnt[64][2];
  const int v = blockIdx.x, wi = blockIdx.y, n_videos = gridDim.x;
  const int n = lens[v], w = windows[wi];
  const double* a = scores + (long long)v * t_max;
  double lo = 1e300, hi = -1e300;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = max(0, i - w), e = min(n, i + w + 1);
    const double m = numpy_pairwise_sum(a + b, e - b) / (double)(e - b);
    sm[i] = m;
    lo = fmin(lo, m);
    hi = fmax(hi, m);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { red_min[threadIdx.x >> 5] = lo; red_max[threadIdx.x >> 5] = hi; }
  for (int t = threadIdx.x; t < n_thr; t += blockDim.x) { cnt[t][0] = 0; cnt[t][1] = 0; }
  __syncthreads();
  lo = red_min[0];
  hi = red_max[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { lo = fmin(lo, red_min[i]); hi = fmax(hi, red_max[i]); }
  const double range = hi - lo;
  if (n == 0) {                                  // the evaluator's max([]) raises here
    if (threadIdx.x == 0) degenerate[wi * n_videos + v] = 1;
    return;
  }
  // range == 0 (a constant smoothed list, e.g. a window wider than the video): the evaluator's np.float64 arithmetic gives
  // 0/0 = nan, every `nan >= threshold` is False, IoU = 0 — the same IEEE operations happen below.
  const uint8_t* g = gold + (long long)v * t_max;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double p = (sm[i] - lo) / range;
    if (norm_out != nullptr) norm_out[((long long)wi * n_videos + v) * t_max + i] = p;
    const bool gd = g[i] != 0;
    for (int t = 0; t < n_thr; ++t) {
      const bool pr = p >= thresholds[t];
      if (pr && gd) atomicAdd(&cnt[t][0], 1);
      if (pr || gd) atomicAdd(&cnt[t][1], 1);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_thr; t += blockDim.x) {
    int* o = counts + (((long long)wi * n_videos + v) * n_thr + t) * 2;
    o[0] = cnt[t][0];
    o[1] = cnt[t][1];
  }
}

}  // namespace

extern "C" int mmd_grounding_sweep(const double* scores, const unsigned char* gold, const int* lens, int n_videos, int t_max,
                                   const int* windows, int n_windows, const double* thresholds, int n_thresholds, int* counts,
                                   double* norm_out, int* degenerate, void* stream) {
  if (scores == nullptr || gold == nullptr || lens == nullptr || windows == nullptr || thresholds == nullptr || counts == nullptr ||
      degenerate == nullptr || n_videos < 0 || t_max <= 0 || n_windows <= 0 || n_thresholds <= 0 || n_thresholds > 64 ||
      (size_t)t_max * sizeof(double) > 200 * 1024)
    return MMD_ERR_ARG;
  if (n_videos == 0) return 0;
  const size_t smem = (size_t)t_max * sizeof(double);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(grounding_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return MMD_ERR_CUDA;
  grounding_sweep_kernel<<<dim3(n_videos, n_windows), 128, smem, static_cast<cudaStream_t>(stream)>>>(
      scores, gold, lens, t_max, windows, thresholds, n_thresholds, counts, norm_out, degenerate);
  return cudaGetLastError() == cudaSuccess ? 0 : MMD_ERR_CUDA;
}
