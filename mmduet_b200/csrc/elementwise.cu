// CUDA-core kernels around the tcgen05 GEMMs: patch im2col (+ pixel normalisation), LayerNorm / RMSNorm with the fp32
// residual stream, split-K reduction fused into the consumers (QKV bias + RoPE + paged-KV append; residual add +
// RMSNorm), row gathers, tap pooling (bilinear / average / max), final norm + score heads, argmax.
// All of them are HBM/latency-bound: 16-B vectorised, coalesced accesses, warp-shuffle reductions.
#include "kernels.cuh"
#include "launch.cuh"

#include <cuda_bf16.h>
#include <math.h>

namespace mmd {

thread_local bool g_use_pdl = false;

// programmatic dependent launch (see ptx.cuh): let the next kernel start its prologue / weight prefetch, then wait for
// the producers of our inputs.  Both are no-ops for kernels launched without the attribute.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = (l < NT / 32) ? sh[l] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------------------
// im2col for the 14x14/stride-14 patch embedding: A[t*G*G + gy*G + gx, c*P*P + py*P + px] (K padded with zeros).
// Replaces nn.Conv2d in SiglipVisionEmbeddings (TF:models/siglip/modeling_siglip.py:124-130) together with the
// image-processor rescale/normalise ((x/255 - 0.5)/0.5; models/vision_live.py:13, LLaVA SigLipImageProcessor).
// ------------------------------------------------------------------------------------------------------------
template <typename TIn, bool NORMALIZE>
__global__ void im2col_kernel(const TIn* __restrict__ px, __nv_bfloat16* __restrict__ A, int T, int C, int img, int P,
                              int G, int Kpad) {
  // one block per (frame, patch-row gy); threads sweep the [C, P, img] strip so global reads are row-contiguous
  const int t = blockIdx.x / G, gy = blockIdx.x % G;
  const int strip = C * P * img;
  for (int e = threadIdx.x; e < strip; e += blockDim.x) {
    const int x = e % img;
    const int py = (e / img) % P;
    const int c = e / (img * P);
    const int gx = x / P, pxx = x % P;
    if (gx >= G) continue;
    float v = static_cast<float>(px[(((long long)t * C + c) * img + (gy * P + py)) * img + x]);
    if (NORMALIZE) v = (v * 0.00392156862745098f - 0.5f) / 0.5f;
    A[((long long)(t * G + gy) * G + gx) * Kpad + (c * P + py) * P + pxx] = __float2bfloat16_rn(v);
  }
  const int Kreal = C * P * P;
  for (int e = threadIdx.x; e < G * (Kpad - Kreal); e += blockDim.x) {
    const int gx = e / (Kpad - Kreal), k = Kreal + e % (Kpad - Kreal);
    A[((long long)(t * G + gy) * G + gx) * Kpad + k] = __float2bfloat16_rn(0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Frame ingest: aspect-preserving 8-bit bilinear resize + centred zero pad + BGR->RGB + HWC->CHW, bit-exact with the
// cv2.resize / copyMakeBorder / cvtColor / transpose chain of the reference (test/datasets.py:50-72).  OpenCV's 8-bit
// INTER_LINEAR is a separable fixed-point filter: weights rounded to 11 bits, horizontal pass into int32, vertical pass
// (b * (S >> 4)) >> 16 summed, + 2, >> 2.  The coordinate arithmetic below uses the same double / float operation
// sequence (explicit _rn intrinsics: no FMA contraction), so the weights are the same integers.
// ---------------------------------------------------------------------------------------------------------------
struct ResizeTap { int i0, i1, w0, w1; };

__device__ __forceinline__ ResizeTap resize_tap(int d, int src, int dst, bool drop_fraction_at_edges) {
  const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dst, (double)src));
  float f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
  int s = __float2int_rd(f);
  f = __fsub_rn(f, (float)s);
  ResizeTap t;
  if (drop_fraction_at_edges) {            // horizontal: clamp the index and zero the fraction (resize.cpp dx loop)
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
    t.i0 = s; t.i1 = min(s + 1, src - 1);
  } else {                                 // vertical: rows are clamped, the fraction is kept
    t.i0 = min(max(s, 0), src - 1); t.i1 = min(max(s + 1, 0), src - 1);
  }
  t.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return t;
}

__global__ void frame_ingest_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, int H, int W, int res, int nw, int nh,
                                    int left, int top) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, t = blockIdx.z;
  if (x4 >= res) return;
  const uint8_t* src = frames + (size_t)t * H * W * 3;
  uint32_t px[3] = {0u, 0u, 0u};           // four output pixels per plane, packed little-endian
  const int yy = y - top;
  if (yy >= 0 && yy < nh) {
    const ResizeTap ty = resize_tap(yy, H, nh, false);
    const uint8_t *r0 = src + (size_t)ty.i0 * W * 3, *r1 = src + (size_t)ty.i1 * W * 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = x4 + k - left;
      if (xx < 0 || xx >= nw) continue;
      const ResizeTap tx = resize_tap(xx, W, nw, true);
#pragma unroll
      for (int c = 0; c < 3; ++c) {        // c indexes the SOURCE (BGR) channel; it lands in plane 2 - c
        const int S0 = r0[tx.i0 * 3 + c] * tx.w0 + r0[tx.i1 * 3 + c] * tx.w1;
        const int S1 = r1[tx.i0 * 3 + c] * tx.w0 + r1[tx.i1 * 3 + c] * tx.w1;
        int v = (((ty.w0 * (S0 >> 4)) >> 16) + ((ty.w1 * (S1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        px[2 - c] |= (uint32_t)v << (8 * k);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
    *reinterpret_cast<uint32_t*>(out + (((size_t)t * 3 + c) * res + y) * res + x4) = px[c];
}

int launch_frame_ingest(const uint8_t* frames, int T, int H, int W, uint8_t* out, int res, cudaStream_t s) {
  const int nw = W > H ? res : (int)(((double)W / (double)H) * res);     // test/datasets.py:51-58
  const int nh = W > H ? (int)(((double)H / (double)W) * res) : res;
  if (nw < 1 || nh < 1 || res % 4 != 0) return -2;
  const int left = (res - nw) / 2, top = (res - nh) / 2;
  dim3 block(96), grid((res / 4 + 95) / 96, res, T);
  frame_ingest_kernel<<<grid, block, 0, s>>>(frames, out, H, W, res, nw, nh, left, top);
  return 0;
}

int launch_im2col(const void* px, int px_dtype, int normalize, __nv_bfloat16* A, int T, int C, int img, int P, int Kpad,
                  cudaStream_t s) {
  const int G = img / P;
  dim3 grid(T * G), block(256);
  if (px_dtype == DT_U8) {
    if (normalize) im2col_kernel<uint8_t, true><<<grid, block, 0, s>>>((const uint8_t*)px, A, T, C, img, P, G, Kpad);
    else im2col_kernel<uint8_t, false><<<grid, block, 0, s>>>((const uint8_t*)px, A, T, C, img, P, G, Kpad);
  } else if (px_dtype == DT_F32) {
    if (normalize) im2col_kernel<float, true><<<grid, block, 0, s>>>((const float*)px, A, T, C, img, P, G, Kpad);
    else im2col_kernel<float, false><<<grid, block, 0, s>>>((const float*)px, A, T, C, img, P, G, Kpad);
  } else if (px_dtype == DT_BF16) {
    if (normalize) im2col_kernel<__nv_bfloat16, true><<<grid, block, 0, s>>>((const __nv_bfloat16*)px, A, T, C, img, P, G, Kpad);
    else im2col_kernel<__nv_bfloat16, false><<<grid, block, 0, s>>>((const __nv_bfloat16*)px, A, T, C, img, P, G, Kpad);
  } else {
    return -2;
  }
  return 0;
}

// residual[row, :] = pos_emb[row % S, :]  (the patch-embed GEMM then accumulates conv + bias on top: EPI_RESID_F32)
__global__ void broadcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int S, int D4) {
  const long long n = rows * D4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D4;
    const int c = (int)(i % D4);
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + (r % S) * D4 + c);
  }
}
int launch_broadcast_rows(const float* src, float* dst, long long rows, int S, int D, cudaStream_t s) {
  if (D % 4) return -2;
  const long long n = rows * (D / 4);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  broadcast_rows_kernel<<<blocks, 256, 0, s>>>(src, dst, rows, S, D / 4);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm: fp32 residual row -> bf16 GEMM operand (or fp32).  One warp per row, row cached in registers.
// Restates nn.LayerNorm(eps=1e-6) of SiglipEncoderLayer (TF:models/siglip/modeling_siglip.py:333-362).
// ------------------------------------------------------------------------------------------------------------
template <int MAXV, bool OUT_F32, bool PLANES>
__global__ void layernorm_kernel(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 void* __restrict__ out, long long rows, int D, float eps, const float* __restrict__ planes,
                                 int n_planes, long long plane_stride, const float* __restrict__ add_bias) {
  // Optional prologue (small-batch ViT): x[row] += add_bias + sum_s planes[s][row] — the deterministic reduction of the
  // previous split-K GEMM (out_proj / fc2) and its residual add — written back before normalising.  gamma == nullptr:
  // only that update.
  pdl_prologue();
  const int warps_per_block = blockDim.x >> 5;
  const long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  float4* xr = reinterpret_cast<float4*>(x + row * D);
  float4 v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int c = lane + j * 32;
    if (c < D4) {
      v[j] = xr[c];
      if constexpr (PLANES) {
        if (add_bias != nullptr) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(add_bias) + c);
          v[j].x += b.x; v[j].y += b.y; v[j].z += b.z; v[j].w += b.w;
        }
#pragma unroll 4
        for (int p = 0; p < n_planes; ++p) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(planes + p * plane_stride + row * D) + c);
          v[j].x += a.x; v[j].y += a.y; v[j].z += a.z; v[j].w += a.w;
        }
        xr[c] = v[j];
      }
      sum += v[j].x + v[j].y + v[j].z + v[j].w;
    }
  }
  if (gamma == nullptr) return;
  const float mean = warp_sum(sum) / D;
  float var = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int c = lane + j * 32;
    if (c < D4) {
      const float a = v[j].x - mean, b = v[j].y - mean, cc = v[j].z - mean, d = v[j].w - mean;
      var += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / D + eps);
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int c = lane + j * 32;
    if (c < D4) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      const float o0 = (v[j].x - mean) * rstd * g.x + b.x, o1 = (v[j].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[j].z - mean) * rstd * g.z + b.z, o3 = (v[j].w - mean) * rstd * g.w + b.w;
      if (OUT_F32) {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * D)[c] = make_float4(o0, o1, o2, o3);
      } else {
        uint2 o;
        o.x = pack2(o0, o1);
        o.y = pack2(o2, o3);
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * D)[c] = o;
      }
    }
  }
}
// Small-batch variant (a few hundred rows): one BLOCK per row, one float4 per thread, so the 1 + n_planes loads of a thread
// are all in flight at once and a 729-row frame still spreads over every SM.
template <int NT>
__global__ void resid_add_layernorm_block_kernel(float* __restrict__ x, const float* __restrict__ planes, int n_planes,
                                                 long long plane_stride, const float* __restrict__ add_bias,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 __nv_bfloat16* __restrict__ out, int D, float eps) {
  __shared__ float red[NT / 32];
  pdl_prologue();
  const long long row = blockIdx.x;
  const int c = threadIdx.x, D4 = D >> 2;
  const bool act = c < D4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) {
    v = reinterpret_cast<float4*>(x + row * D)[c];
    if (add_bias != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(add_bias) + c);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
#pragma unroll 8
    for (int p = 0; p < n_planes; ++p) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(planes + p * plane_stride + row * D) + c);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    reinterpret_cast<float4*>(x + row * D)[c] = v;
  }
  if (gamma == nullptr) return;
  const float mean = block_sum<NT>(act ? v.x + v.y + v.z + v.w : 0.f, red) / D;
  const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
  const float rstd = rsqrtf(block_sum<NT>(act ? a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 : 0.f, red) / D + eps);
  if (act) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
    uint2 o;
    o.x = pack2(a0 * rstd * g.x + b.x, a1 * rstd * g.y + b.y);
    o.y = pack2(a2 * rstd * g.z + b.z, a3 * rstd * g.w + b.w);
    reinterpret_cast<uint2*>(out + row * D)[c] = o;
  }
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_f32, long long rows, int D,
                     float eps, cudaStream_t s) {
  return launch_resid_add_layernorm(const_cast<float*>(x), nullptr, 0, 0, nullptr, gamma, beta, out, out_f32, rows, D, eps, s);
}
int launch_resid_add_layernorm(float* x, const float* planes, int n_planes, long long plane_stride, const float* add_bias,
                               const float* gamma, const float* beta, void* out, int out_f32, long long rows, int D, float eps,
                               cudaStream_t s) {
  if (D % 4 != 0 || D > 12 * 128) return -2;
  if (rows <= 0) return 0;
  const int wpb = 8;
  const unsigned blocks = (unsigned)((rows + wpb - 1) / wpb);
  if (n_planes > 0) {   // small-batch ViT path: bf16 output (or none) only
    if (out_f32) return -2;
    if (D <= 4 * 384)
      launch_k(resid_add_layernorm_block_kernel<384>, dim3((unsigned)rows), dim3(384), 0, s, x, planes, n_planes, plane_stride, add_bias,
               gamma, beta, static_cast<__nv_bfloat16*>(out), D, eps);
    else
      launch_k(layernorm_kernel<12, false, true>, dim3(blocks), dim3(wpb * 32), 0, s, x, gamma, beta, out, rows, D, eps, planes, n_planes,
               plane_stride, add_bias);
  } else if (gamma == nullptr) {
    return 0;           // nothing to add, nothing to normalise
  } else if (out_f32) {
    launch_k(layernorm_kernel<12, true, false>, dim3(blocks), dim3(wpb * 32), 0, s, x, gamma, beta, out, rows, D, eps,
             (const float*)nullptr, 0, 0ll, (const float*)nullptr);
  } else {
    launch_k(layernorm_kernel<12, false, false>, dim3(blocks), dim3(wpb * 32), 0, s, x, gamma, beta, out, rows, D, eps,
             (const float*)nullptr, 0, 0ll, (const float*)nullptr);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// resid[row,:] += sum_s partial[s][row,:]  (split-K planes of o_proj / down_proj), then RMSNorm -> bf16 operand of the
// next GEMM (and optionally the fp32 normalised row).  Qwen2RMSNorm: w * (x * rsqrt(mean(x^2) + eps))
// (TF:models/qwen2/modeling_qwen2.py:249-263); residual adds of Qwen2DecoderLayer (:280-310).
// ------------------------------------------------------------------------------------------------------------
// "Precise rows" (DESIGN.md §2): rows listed in `prec_of_row` (row -> index j into the precise buffers, or -1) take their
// residual update from `prec_partial` (the hi/lo side GEMM's planes [n_prec_planes][P][H]) instead of `partial`, and get a
// second, [hi | lo] (2H wide) copy of the normalised row in `out_hilo[j]`.  prec_of_row == nullptr with out_hilo != nullptr
// means every row is precise (j = row) and `partial` already is the precise result (single-frame steps).
constexpr int kMaxPlanes = 8;   // split-K planes whose loads are issued together (more are added in a plain loop)

struct PreciseRows {
  const int* prec_of_row = nullptr;
  const float* prec_partial = nullptr;
  int n_prec_planes = 0;
  long long prec_plane_stride = 0;
  __nv_bfloat16* out_hilo = nullptr;
  // "appended rows" form (passes above 128 tokens): the precise rows travel through the MAIN GEMMs as 2P extra activation rows,
  // [n_rows + j] = hi and [n_rows + P + j] = lo of precise row j, so the weights are streamed once.  On the way in, the
  // residual update of precise row j is prec_partial[j] + prec_partial[j + pair_offset / H] (the hi and lo rows' outputs);
  // on the way out, the normalised row is also written to out_bf16 rows n_rows + j (hi) and n_rows + n_prec + j (lo).
  long long pair_offset = 0;     // elements between the hi and the lo output row inside a plane (0: single row)
  int append_base = -1;          // n_rows of the main pass (>= 0 switches the appended output rows on)
  int n_prec = 0;
};

template <int NT>
__global__ void resid_add_rmsnorm_kernel(float* __restrict__ resid, const float* __restrict__ partial, int n_planes,
                                         long long plane_stride, const float* __restrict__ w, __nv_bfloat16* __restrict__ out_bf16,
                                         float* __restrict__ out_f32, int H, float eps, PreciseRows pr) {
  extern __shared__ float row_sh[];  // H floats
  __shared__ float red[NT / 32];
  pdl_prologue();
  const long long row = blockIdx.x;
  const int H4 = H >> 2;
  long long j = -1;
  if (pr.prec_of_row != nullptr) j = pr.prec_of_row[row];
  else if (pr.out_hilo != nullptr) j = row;
  const float* src = partial + row * H;
  int np = n_planes;
  long long ps = plane_stride;
  long long pair = 0;
  if (j >= 0 && pr.prec_partial != nullptr) { src = pr.prec_partial + j * H; np = pr.n_prec_planes; ps = pr.prec_plane_stride; pair = pr.pair_offset; }
  float ss = 0.f;
  for (int c = threadIdx.x; c < H4; c += NT) {
    float4 v = reinterpret_cast<float4*>(resid + row * H)[c];
    // all plane loads of this chunk are issued before the first add (the step is latency-bound: 49 rows, <= 8 planes in L2)
    float4 a[kMaxPlanes];
#pragma unroll
    for (int p = 0; p < kMaxPlanes; ++p)
      a[p] = p < np ? __ldg(reinterpret_cast<const float4*>(src + p * ps) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kMaxPlanes; ++p) { v.x += a[p].x; v.y += a[p].y; v.z += a[p].z; v.w += a[p].w; }
    if (pair != 0) {             // the lo copy's output row of a precise row (appended-rows form)
#pragma unroll
      for (int p = 0; p < kMaxPlanes; ++p)
        a[p] = p < np ? __ldg(reinterpret_cast<const float4*>(src + pair + p * ps) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kMaxPlanes; ++p) { v.x += a[p].x; v.y += a[p].y; v.z += a[p].z; v.w += a[p].w; }
    }
    for (int p = kMaxPlanes; p < np; ++p) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + p * ps) + c);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (np > 0) reinterpret_cast<float4*>(resid + row * H)[c] = v;
    reinterpret_cast<float4*>(row_sh)[c] = v;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float tot = block_sum<NT>(ss, red);
  if (w == nullptr) return;
  const float rstd = rsqrtf(tot / H + eps);
  for (int c = threadIdx.x; c < H4; c += NT) {
    const float4 v = reinterpret_cast<float4*>(row_sh)[c];
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + c);
    const float o0 = v.x * rstd * g.x, o1 = v.y * rstd * g.y, o2 = v.z * rstd * g.z, o3 = v.w * rstd * g.w;
    uint2 o;
    o.x = pack2(o0, o1);
    o.y = pack2(o2, o3);
    if (out_bf16 != nullptr) reinterpret_cast<uint2*>(out_bf16 + row * H)[c] = o;
    if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32 + row * H)[c] = make_float4(o0, o1, o2, o3);
    if (j >= 0 && (pr.out_hilo != nullptr || pr.append_base >= 0)) {
      const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o.x));
      const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o.y));
      uint2 l;
      l.x = pack2(o0 - h0.x, o1 - h0.y);
      l.y = pack2(o2 - h1.x, o3 - h1.y);
      if (pr.out_hilo != nullptr) {
        uint2* d = reinterpret_cast<uint2*>(pr.out_hilo + j * 2 * H);
        d[c] = o;
        d[H4 + c] = l;
      }
      if (pr.append_base >= 0 && out_bf16 != nullptr) {
        reinterpret_cast<uint2*>(out_bf16 + (pr.append_base + j) * H)[c] = o;
        reinterpret_cast<uint2*>(out_bf16 + (pr.append_base + pr.n_prec + j) * H)[c] = l;
      }
    }
  }
}
int launch_resid_add_rmsnorm(float* resid, const float* partial, int n_planes, long long plane_stride, const float* w,
                             __nv_bfloat16* out_bf16, float* out_f32, long long rows, int H, float eps, cudaStream_t s) {
  return launch_resid_add_rmsnorm_precise(resid, partial, n_planes, plane_stride, w, out_bf16, out_f32, rows, H, eps, nullptr, nullptr, 0,
                                          0, nullptr, s, 0, 0, 0);
}
int launch_resid_add_rmsnorm_precise(float* resid, const float* partial, int n_planes, long long plane_stride, const float* w,
                                     __nv_bfloat16* out_bf16, float* out_f32, long long rows, int H, float eps, const int* prec_of_row,
                                     const float* prec_partial, int n_prec_planes, long long prec_plane_stride, __nv_bfloat16* out_hilo,
                                     cudaStream_t s, long long pair_offset, int append_rows, int n_prec) {
  if (H % 4 != 0 || rows <= 0) return rows == 0 ? 0 : -2;
  PreciseRows pr;
  pr.prec_of_row = prec_of_row; pr.prec_partial = prec_partial; pr.n_prec_planes = n_prec_planes;
  pr.prec_plane_stride = prec_plane_stride; pr.out_hilo = out_hilo;
  pr.pair_offset = pair_offset; pr.append_base = append_rows ? (int)rows : -1; pr.n_prec = n_prec;
  if (rows <= 256 && H >= 2048)   // few rows (single-frame steps): one float4 chunk or two per thread, every load in flight at once
    launch_k(resid_add_rmsnorm_kernel<512>, dim3((unsigned)rows), dim3(512), H * sizeof(float), s, resid, partial, n_planes,
             plane_stride, w, out_bf16, out_f32, H, eps, pr);
  else
    launch_k(resid_add_rmsnorm_kernel<256>, dim3((unsigned)rows), dim3(256), H * sizeof(float), s, resid, partial, n_planes,
             plane_stride, w, out_bf16, out_f32, H, eps, pr);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Final RMSNorm with the score heads as its epilogue, on the rows that are read only: for each of the `n_score` score rows
//   r = resid[row] + sum planes (the last down_proj),  n = w * r * rsqrt(mean r^2 + eps)   (model.norm)
//   logits = n . {informative_head[0], [1], relevance_head[0], [1]}  (fp32),  score = softmax(.)[1] = sigmoid(l1 - l0)
// (video_head_live_llava_qwen.py:152-161; test/inference.py:243-244), and for each of the `n_lm` lm rows the normalised row as
// the bf16 operand of lm_head.  No [M, H] hidden-state buffer is written and the other M - n rows are not normalised at all.
// ------------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void final_norm_heads_kernel(const float* __restrict__ resid, const float* __restrict__ partial, int n_planes,
                                        long long plane_stride, const float* __restrict__ w, const int* __restrict__ score_rows,
                                        int n_score, const int* __restrict__ lm_rows, const float* __restrict__ head_w,
                                        float* __restrict__ logits_out, float* __restrict__ scores_out,
                                        __nv_bfloat16* __restrict__ lm_x, int H, float eps, PreciseRows pr) {
  extern __shared__ float row_sh[];  // H floats
  __shared__ float red[NT / 32];
  __shared__ float hred[4][NT / 32];
  pdl_prologue();
  const int b = blockIdx.x;
  const bool is_score = b < n_score;
  const long long row = is_score ? score_rows[b] : lm_rows[b - n_score];
  const int H4 = H >> 2;
  long long j = pr.prec_of_row != nullptr ? pr.prec_of_row[row] : -1;
  const float* src = partial + row * H;
  int np = n_planes;
  long long ps = plane_stride;
  long long pair = 0;
  if (j >= 0 && pr.prec_partial != nullptr) { src = pr.prec_partial + j * H; np = pr.n_prec_planes; ps = pr.prec_plane_stride; pair = pr.pair_offset; }
  float ss = 0.f;
  for (int c = threadIdx.x; c < H4; c += NT) {
    float4 v = reinterpret_cast<const float4*>(resid + row * H)[c];
    float4 a[kMaxPlanes];
#pragma unroll
    for (int p = 0; p < kMaxPlanes; ++p)
      a[p] = p < np ? __ldg(reinterpret_cast<const float4*>(src + p * ps) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kMaxPlanes; ++p) { v.x += a[p].x; v.y += a[p].y; v.z += a[p].z; v.w += a[p].w; }
    if (pair != 0) {             // the lo copy's output row of a precise row (appended-rows form)
#pragma unroll
      for (int p = 0; p < kMaxPlanes; ++p)
        a[p] = p < np ? __ldg(reinterpret_cast<const float4*>(src + pair + p * ps) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kMaxPlanes; ++p) { v.x += a[p].x; v.y += a[p].y; v.z += a[p].z; v.w += a[p].w; }
    }
    for (int p = kMaxPlanes; p < np; ++p) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + p * ps) + c);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    reinterpret_cast<float4*>(row_sh)[c] = v;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float rstd = rsqrtf(block_sum<NT>(ss, red) / H + eps);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = threadIdx.x; c < H4; c += NT) {
    const float4 v = reinterpret_cast<float4*>(row_sh)[c];
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + c);
    const float o0 = v.x * rstd * g.x, o1 = v.y * rstd * g.y, o2 = v.z * rstd * g.z, o3 = v.w * rstd * g.w;
    if (is_score) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 hw = __ldg(reinterpret_cast<const float4*>(head_w + (long long)k * H) + c);
        acc[k] = fmaf(o0, hw.x, fmaf(o1, hw.y, fmaf(o2, hw.z, fmaf(o3, hw.w, acc[k]))));
      }
    } else {
      uint2 o;
      o.x = pack2(o0, o1);
      o.y = pack2(o2, o3);
      reinterpret_cast<uint2*>(lm_x + (long long)(b - n_score) * H)[c] = o;
    }
  }
  if (!is_score) return;
  const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc[k] = warp_sum(acc[k]);
    if (l == 0) hred[k][wi] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float lg[4];
    for (int k = 0; k < 4; ++k) {
      float sacc = 0.f;
      for (int i = 0; i < NT / 32; ++i) sacc += hred[k][i];
      lg[k] = sacc;
      logits_out[b * 4 + k] = sacc;
    }
    scores_out[b * 2 + 0] = 1.f / (1.f + expf(lg[0] - lg[1]));
    scores_out[b * 2 + 1] = 1.f / (1.f + expf(lg[2] - lg[3]));
  }
}
int launch_final_norm_heads(const float* resid, const float* partial, int n_planes, long long plane_stride, const float* w,
                            const int* score_rows, int n_score, const int* lm_rows, int n_lm, const float* head_w, float* logits_out,
                            float* scores_out, __nv_bfloat16* lm_x, int H, float eps, const int* prec_of_row, const float* prec_partial,
                            int n_prec_planes, long long prec_plane_stride, cudaStream_t s, long long pair_offset) {
  if (H % 4 != 0 || n_score < 0 || n_lm < 0) return -2;
  if (n_score + n_lm == 0) return 0;
  if ((n_score > 0 && (score_rows == nullptr || head_w == nullptr || logits_out == nullptr || scores_out == nullptr)) ||
      (n_lm > 0 && (lm_rows == nullptr || lm_x == nullptr)))
    return -2;
  constexpr int NT = 256;
  PreciseRows pr;
  pr.prec_of_row = prec_of_row; pr.prec_partial = prec_partial; pr.n_prec_planes = n_prec_planes; pr.prec_plane_stride = prec_plane_stride;
  pr.pair_offset = pair_offset;
  launch_k(final_norm_heads_kernel<NT>, dim3((unsigned)(n_score + n_lm)), dim3(NT), H * sizeof(float), s, resid, partial, n_planes,
           plane_stride, w, score_rows, n_score, lm_rows, head_w, logits_out, scores_out, lm_x, H, eps, pr);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// SwiGLU of the precise rows from the raw pre-activations their hi and lo copies produced in the main gate/up GEMM:
//   g = raw[j][2i] + raw[P + j][2i],  u = raw[j][2i+1] + raw[P + j][2i+1],  h = silu(g) * u
// written as two appended activation rows of the down projection: h_rows[j] = bf16(h), h_rows[P + j] = bf16(h - bf16(h)).
// ------------------------------------------------------------------------------------------------------------
__global__ void swiglu_from_raw_kernel(const float* __restrict__ raw, long long ld_raw, __nv_bfloat16* __restrict__ h_rows, long long ldh,
                                       int P, int I) {
  pdl_prologue();
  const int j = blockIdx.y;
  const float* hi = raw + (long long)j * ld_raw;
  const float* lo = raw + (long long)(P + j) * ld_raw;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2; i < I; i += gridDim.x * blockDim.x * 2) {
    const float4 a = *reinterpret_cast<const float4*>(hi + 2 * i), b = *reinterpret_cast<const float4*>(lo + 2 * i);
    const float g0 = a.x + b.x, u0 = a.y + b.y, g1 = a.z + b.z, u1 = a.w + b.w;
    const float h0 = g0 / (1.0f + __expf(-g0)) * u0, h1 = g1 / (1.0f + __expf(-g1)) * u1;
    const uint32_t hp = pack2(h0, h1);
    const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hp));
    *reinterpret_cast<uint32_t*>(h_rows + (long long)j * ldh + i) = hp;
    *reinterpret_cast<uint32_t*>(h_rows + (long long)(P + j) * ldh + i) = pack2(h0 - hf.x, h1 - hf.y);
  }
}
int launch_swiglu_from_raw(const float* raw, long long ld_raw, __nv_bfloat16* h_rows, long long ldh, int P, int I, cudaStream_t s) {
  if (P <= 0) return 0;
  if (I % 2 != 0 || ld_raw % 4 != 0 || ldh % 2 != 0) return -2;
  int bx = (I / 2 + 255) / 256;
  if (bx > 16) bx = 16;
  launch_k(swiglu_from_raw_kernel, dim3(bx, P), dim3(256), 0, s, raw, ld_raw, h_rows, ldh, P, I);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// QKV finish: sum split-K planes + bias, rotate q/k (RoPE, rotate_half convention), write Q as bf16 [M, Hq, dh] and
// append K/V (bf16) to the paged KV pool at the token's physical slot.  Replaces q/k/v bias add, apply_rotary_pos_emb
// and DynamicCache.update's torch.cat (TF:models/qwen2/modeling_qwen2.py:127-146,215-233; TF:cache_utils.py:119-120).
// KV pool layout per layer: [page][2 (K,V)][kv_head][PAGE_TOKENS][dh]; slot = page * PAGE_TOKENS + offset.
// ------------------------------------------------------------------------------------------------------------
__global__ void qkv_finish_kernel(const float* __restrict__ partial, int n_planes, long long plane_stride,
                                  const float* __restrict__ bias, const float* __restrict__ cos_tab,
                                  const float* __restrict__ sin_tab, const int* __restrict__ tok_pos,
                                  const int* __restrict__ tok_slot, __nv_bfloat16* __restrict__ q_out,
                                  __nv_bfloat16* __restrict__ kv_layer, int Hq, int Hkv, int dh, int page_tokens,
                                  const int* __restrict__ prec_of_row, int n_prec) {
  // one block per token; 8 threads per head, thread t handles the 8 rotation pairs (d, d + dh/2) with d in [8t', 8t'+8)
  // (dh = 128: half = 64 = 8 threads x 8 elements): 32-B vector loads of the split-K planes, 16-B bf16 stores
  pdl_prologue();
  const int tok = blockIdx.x;
  const int head = threadIdx.x >> 3;
  const int d0 = (threadIdx.x & 7) * 8;
  const int half = dh >> 1;
  const int N = (Hq + 2 * Hkv) * dh;
  const int col0 = head * dh + d0, col1 = col0 + half;
  float x0[8], x1[8];
  if (bias != nullptr) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(bias + col0)), b = __ldg(reinterpret_cast<const float4*>(bias + col0 + 4));
    const float4 c = __ldg(reinterpret_cast<const float4*>(bias + col1)), d = __ldg(reinterpret_cast<const float4*>(bias + col1 + 4));
    x0[0] = a.x; x0[1] = a.y; x0[2] = a.z; x0[3] = a.w; x0[4] = b.x; x0[5] = b.y; x0[6] = b.z; x0[7] = b.w;
    x1[0] = c.x; x1[1] = c.y; x1[2] = c.z; x1[3] = c.w; x1[4] = d.x; x1[5] = d.y; x1[6] = d.z; x1[7] = d.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) { x0[i] = 0.f; x1[i] = 0.f; }
  }
  // precise rows (appended-rows form): the token's projections are the sum of its hi copy's and its lo copy's output rows
  // (rows M + j and M + P + j of every plane) instead of its own row
  const int pj = prec_of_row != nullptr ? prec_of_row[tok] : -1;
  const int n_src = pj >= 0 ? 2 : 1;
  for (int srcs = 0; srcs < n_src; ++srcs)
  for (int p0 = 0; p0 < n_planes; p0 += 4) {      // four planes' loads (16 x 16 B per thread) in flight together
    float4 a[4], b[4], c[4], d[4];
    const long long row = pj >= 0 ? (long long)gridDim.x + pj + (srcs ? n_prec : 0) : tok;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool on = p0 + i < n_planes;
      const float* pl = partial + (long long)(on ? p0 + i : 0) * plane_stride + row * N;
      a[i] = on ? __ldg(reinterpret_cast<const float4*>(pl + col0)) : z;
      b[i] = on ? __ldg(reinterpret_cast<const float4*>(pl + col0 + 4)) : z;
      c[i] = on ? __ldg(reinterpret_cast<const float4*>(pl + col1)) : z;
      d[i] = on ? __ldg(reinterpret_cast<const float4*>(pl + col1 + 4)) : z;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x0[0] += a[i].x; x0[1] += a[i].y; x0[2] += a[i].z; x0[3] += a[i].w; x0[4] += b[i].x; x0[5] += b[i].y; x0[6] += b[i].z; x0[7] += b[i].w;
      x1[0] += c[i].x; x1[1] += c[i].y; x1[2] += c[i].z; x1[3] += c[i].w; x1[4] += d[i].x; x1[5] += d[i].y; x1[6] += d[i].z; x1[7] += d[i].w;
    }
  }
  if (head < Hq + Hkv) {  // q and k are rotated (rotate_half convention)
    const int pos = tok_pos[tok];
    const float* cp = cos_tab + (long long)pos * half + d0;
    const float* sp = sin_tab + (long long)pos * half + d0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float c = __ldg(cp + i), sn = __ldg(sp + i);
      const float r0 = x0[i] * c - x1[i] * sn;
      const float r1 = x1[i] * c + x0[i] * sn;
      x0[i] = r0;
      x1[i] = r1;
    }
  }
  uint4 o0, o1;
  o0.x = pack2(x0[0], x0[1]); o0.y = pack2(x0[2], x0[3]); o0.z = pack2(x0[4], x0[5]); o0.w = pack2(x0[6], x0[7]);
  o1.x = pack2(x1[0], x1[1]); o1.y = pack2(x1[2], x1[3]); o1.z = pack2(x1[4], x1[5]); o1.w = pack2(x1[6], x1[7]);
  __nv_bfloat16* dst;
  if (head < Hq) {
    dst = q_out + ((long long)tok * Hq + head) * dh;
  } else {
    const int is_v = head >= Hq + Hkv;
    const int kvh = head - Hq - (is_v ? Hkv : 0);
    const int slot = tok_slot[tok];
    const int page = slot / page_tokens, off = slot % page_tokens;
    dst = kv_layer + ((((long long)page * 2 + is_v) * Hkv + kvh) * page_tokens + off) * dh;
  }
  *reinterpret_cast<uint4*>(dst + d0) = o0;
  *reinterpret_cast<uint4*>(dst + half + d0) = o1;
}
int launch_qkv_finish(const float* partial, int n_planes, long long plane_stride, const float* bias, const float* cos_tab,
                      const float* sin_tab, const int* tok_pos, const int* tok_slot, __nv_bfloat16* q_out,
                      __nv_bfloat16* kv_layer, int M, int Hq, int Hkv, int dh, int page_tokens, cudaStream_t s, const int* prec_of_row,
                      int n_prec) {
  if (M <= 0) return 0;
  const int heads = Hq + 2 * Hkv;
  if (dh != 128 || heads * 8 > 1024) return -2;
  launch_k(qkv_finish_kernel, dim3(M), dim3(heads * 8), 0, s, partial, n_planes, plane_stride, bias, cos_tab, sin_tab, tok_pos, tok_slot,
           q_out, kv_layer, Hq, Hkv, dh, page_tokens, prec_of_row, n_prec);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Row gathers.
//   embed rows (bf16 table) -> fp32 residual rows: get_input_embeddings()(ids) + torch.cat with the frame tokens
//   (test/inference.py:235-238); src_row < 0 means "take row -(src_row+1) of `other`" (the frame-token buffer).
// ------------------------------------------------------------------------------------------------------------
__global__ void gather_rows_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ table, const __nv_bfloat16* __restrict__ other,
                                               const int* __restrict__ src_row, float* __restrict__ dst, int H8) {
  pdl_prologue();
  const long long row = blockIdx.x;
  const int sr = src_row[row];
  const uint4* src = reinterpret_cast<const uint4*>(sr >= 0 ? table + (long long)sr * H8 * 8 : other + (long long)(-(sr + 1)) * H8 * 8);
  float4* d = reinterpret_cast<float4*>(dst + row * H8 * 8);
  for (int c = threadIdx.x; c < H8; c += blockDim.x) {
    const uint4 u = __ldg(src + c);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    const float2 cc = __bfloat1622float2(h[2]), e = __bfloat1622float2(h[3]);
    d[2 * c] = make_float4(a.x, a.y, b.x, b.y);
    d[2 * c + 1] = make_float4(cc.x, cc.y, e.x, e.y);
  }
}
int launch_gather_rows_bf16_to_f32(const __nv_bfloat16* table, const __nv_bfloat16* other, const int* src_row, float* dst,
                                   long long rows, int H, cudaStream_t s) {
  if (H % 8) return -2;
  if (rows <= 0) return 0;
  launch_k(gather_rows_bf16_to_f32_kernel, dim3((unsigned)rows), dim3(128), 0, s, table, other, src_row, dst, H / 8);
  return 0;
}

// fp32 rows (ViT residual stream, pre-post_layernorm) -> bf16 rows, keeping only the tokens the pooling reads:
// dst[t*G + g, :] = bf16(src[t*S + idx[g], :]).  The tower output is cast to the model dtype before mm_projector
// (video_head_live_llava_qwen.py:96-98, :90-91).
// With HILO the destination row is [hi | lo] (2*D wide): hi = bf16(x), lo = bf16(x - hi).
template <bool HILO>
__global__ void gather_rows_f32_to_bf16_kernel(const float* __restrict__ src, const int* __restrict__ idx, __nv_bfloat16* __restrict__ dst,
                                               int S, int G, int D4) {
  const long long orow = blockIdx.x;
  const long long t = orow / G;
  const int g = (int)(orow % G);
  const float4* s4 = reinterpret_cast<const float4*>(src + (t * S + idx[g]) * (long long)D4 * 4);
  uint2* d = reinterpret_cast<uint2*>(dst + orow * (long long)D4 * 4 * (HILO ? 2 : 1));
  for (int c = threadIdx.x; c < D4; c += blockDim.x) {
    const float4 v = s4[c];
    uint2 o;
    o.x = pack2(v.x, v.y);
    o.y = pack2(v.z, v.w);
    d[c] = o;
    if (HILO) {
      const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o.x));
      const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o.y));
      uint2 l;
      l.x = pack2(v.x - h0.x, v.y - h0.y);
      l.y = pack2(v.z - h1.x, v.w - h1.y);
      d[D4 + c] = l;
    }
  }
}
int launch_gather_rows_f32_to_bf16(const float* src, const int* idx, __nv_bfloat16* dst, int T, int S, int G, int D, int hilo, cudaStream_t s) {
  if (D % 4) return -2;
  if (T <= 0) return 0;
  if (hilo) gather_rows_f32_to_bf16_kernel<true><<<(unsigned)(T * G), 128, 0, s>>>(src, idx, dst, S, G, D / 4);
  else gather_rows_f32_to_bf16_kernel<false><<<(unsigned)(T * G), 128, 0, s>>>(src, idx, dst, S, G, D / 4);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Tap pooling: out[t, o, :] = sum_j w[o][j] * in[t, idx[o][j], :]   (bilinear resize / average pool), or the max
// over the taps.  Tap tables are built on the host from torch's own F.interpolate / pooling applied to an identity
// basis, so the weights are exactly the reference's (video_head_live_llava_qwen.py:100-119; vision_live.py:19-25).
// ------------------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut, bool MAXPOOL>
__global__ void tap_pool_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, const int* __restrict__ tap_idx,
                                const float* __restrict__ tap_w, int n_in, int n_out, int max_taps, int D) {
  const long long orow = blockIdx.x;
  const long long t = orow / n_out;
  const int o = (int)(orow % n_out);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = MAXPOOL ? -INFINITY : 0.f;
    for (int j = 0; j < max_taps; ++j) {
      const int src = tap_idx[o * max_taps + j];
      if (src < 0) break;
      const float v = static_cast<float>(in[(t * n_in + src) * (long long)D + c]);
      if (MAXPOOL) acc = fmaxf(acc, v);
      else acc = fmaf(tap_w[o * max_taps + j], v, acc);
    }
    out[orow * (long long)D + c] = static_cast<TOut>(acc);
  }
}
int launch_tap_pool(const void* in, int in_dtype, void* out, int out_dtype, const int* tap_idx, const float* tap_w, int T,
                    int n_in, int n_out, int max_taps, int D, int maxpool, cudaStream_t s) {
  if (T <= 0) return 0;
  dim3 grid(T * n_out), block(256);
#define TAP_LAUNCH(TI, TO)                                                                                              \
  do {                                                                                                                  \
    if (maxpool) tap_pool_kernel<TI, TO, true><<<grid, block, 0, s>>>((const TI*)in, (TO*)out, tap_idx, tap_w, n_in, n_out, max_taps, D); \
    else tap_pool_kernel<TI, TO, false><<<grid, block, 0, s>>>((const TI*)in, (TO*)out, tap_idx, tap_w, n_in, n_out, max_taps, D);        \
  } while (0)
  if (in_dtype == DT_BF16 && out_dtype == DT_BF16) TAP_LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == DT_F32 && out_dtype == DT_F32) TAP_LAUNCH(float, float);
  else if (in_dtype == DT_F32 && out_dtype == DT_BF16) TAP_LAUNCH(float, __nv_bfloat16);
  else if (in_dtype == DT_BF16 && out_dtype == DT_F32) TAP_LAUNCH(__nv_bfloat16, float);
  else return -2;
#undef TAP_LAUNCH
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Score heads on selected rows of the final-norm output: informative_head / relevance_head (nn.Linear(H, 2, bias=False),
// video_head_live_llava_qwen.py:77-78,160-161), .float(), softmax(-1)[1] (test/inference.py:243-244) = sigmoid(l1 - l0).
// logits_out[r] = {inf0, inf1, rel0, rel1}; scores_out[r] = {informative_score, relevance_score}.
// ------------------------------------------------------------------------------------------------------------
__global__ void heads_kernel(const float* __restrict__ hidden_f32, const int* __restrict__ rows, const float* __restrict__ head_w,
                             float* __restrict__ logits_out, float* __restrict__ scores_out, int H) {
  __shared__ float red[4][8];
  pdl_prologue();
  const int r = blockIdx.x;
  const float* h = hidden_f32 + (long long)rows[r] * H;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float v = h[c];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = fmaf(v, __ldg(head_w + (long long)k * H + c), acc[k]);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc[k] = warp_sum(acc[k]);
    if (l == 0) red[k][w] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float lg[4];
    for (int k = 0; k < 4; ++k) {
      float sacc = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) sacc += red[k][i];
      lg[k] = sacc;
      logits_out[r * 4 + k] = sacc;
    }
    scores_out[r * 2 + 0] = 1.f / (1.f + expf(lg[0] - lg[1]));
    scores_out[r * 2 + 1] = 1.f / (1.f + expf(lg[2] - lg[3]));
  }
}
int launch_heads(const float* hidden_f32, const int* rows, const float* head_w, float* logits_out, float* scores_out, int n_rows,
                 int H, cudaStream_t s) {
  if (n_rows <= 0) return 0;
  launch_k(heads_kernel, dim3(n_rows), dim3(256), 0, s, hidden_f32, rows, head_w, logits_out, scores_out, H);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Greedy pick over lm_head logits of one row (split-K planes summed on the fly), with the HF repetition penalty
// (logit < 0 ? logit * p : logit / p for previously generated ids; models/modeling_live.py:51-77).
// ------------------------------------------------------------------------------------------------------------
__global__ void argmax_kernel(const float* __restrict__ partial, int n_planes, long long plane_stride, int V,
                              const long long* __restrict__ penal_ids, int n_penal, float penalty, long long* __restrict__ out_id,
                              float* __restrict__ out_logit) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  extern __shared__ uint32_t penal_bits[];   // one bit per vocabulary entry (only when n_penal > 0)
  if (n_penal > 0) {
    for (int w = threadIdx.x; w < (V + 31) / 32; w += blockDim.x) penal_bits[w] = 0u;
    __syncthreads();
    for (int j = threadIdx.x; j < n_penal; j += blockDim.x) {
      const long long id = penal_ids[j];
      if (id >= 0 && id < V) atomicOr(&penal_bits[id >> 5], 1u << (id & 31));
    }
    __syncthreads();
  }
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    float v = 0.f;
    for (int p = 0; p < n_planes; ++p) v += __ldg(partial + p * plane_stride + c);
    if (n_penal > 0 && ((penal_bits[c >> 5] >> (c & 31)) & 1u)) v = v < 0.f ? v * penalty : v / penalty;
    if (v > best || (v == best && c < besti)) { best = v; besti = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { bv[w] = best; bi[w] = besti; }
  __syncthreads();
  if (w == 0) {
    best = l < (int)(blockDim.x >> 5) ? bv[l] : -INFINITY;
    besti = l < (int)(blockDim.x >> 5) ? bi[l] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (l == 0) {
      *out_id = besti;
      if (out_logit) *out_logit = best;
    }
  }
}
int launch_argmax(const float* partial, int n_planes, long long plane_stride, int V, const long long* penal_ids, int n_penal,
                  float penalty, long long* out_id, float* out_logit, cudaStream_t s) {
  const size_t smem = n_penal > 0 ? (size_t)((V + 31) / 32) * sizeof(uint32_t) : 0;
  if (smem > 48 * 1024) return -2;            // vocabulary above 393k entries: not this model family
  argmax_kernel<<<1, 1024, smem, s>>>(partial, n_planes, plane_stride, V, penal_ids, n_penal, penalty, out_id, out_logit);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// SigLIP attention-pooling head, the attention itself: ONE learned probe query per frame attends over the S patch tokens
// (SiglipMultiheadAttentionPoolingHead -> nn.MultiheadAttention, TF:models/siglip/modeling_siglip.py; the CLS token of
// models/vision_live.py:26-30).  q [H*dh] fp32 is the projected probe, already scaled by dh^-0.5 (a constant of the
// weights); kv bf16 [T*S, 2*H*dh] = [K | V] from the in-projection GEMM.  One block per (frame, head): scores and softmax
// in fp32 shared memory, V accumulation split over the warps.  out fp32 [T, H*dh].
// ------------------------------------------------------------------------------------------------------------
__global__ void probe_attention_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kv, float* __restrict__ out,
                                       int S, int H, int dh) {
  extern __shared__ float psm[];                 // S scores, then (warps x dh) partial outputs
  __shared__ float red[32];
  const int t = blockIdx.x, h = blockIdx.y, D = H * dh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const __nv_bfloat16* kbase = kv + (long long)t * S * 2 * D + h * dh;
  const float* qh = q + h * dh;
  float mx = -INFINITY;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const __nv_bfloat16* kr = kbase + (long long)s * 2 * D;
    float acc = 0.f;
    for (int d = 0; d < dh; d += 2) {
      const float2 kf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(kr + d));
      acc = fmaf(kf.x, qh[d], fmaf(kf.y, qh[d + 1], acc));
    }
    psm[s] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < nw; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float p = expf(psm[s] - mx);
    psm[s] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < nw; ++i) sum += red[i];
  // V: warp w takes keys w, w + nw, ...; lane covers dims lane, lane + 32, ... (dh <= 128)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = warp; s < S; s += nw) {
    const float p = psm[s];
    const __nv_bfloat16* vr = kbase + (long long)s * 2 * D + D;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = lane + 32 * i;
      if (d < dh) acc[i] = fmaf(p, __bfloat162float(vr[d]), acc[i]);
    }
  }
  float* part = psm + S;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d < dh) part[warp * dh + d] = acc[i];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < dh; d += blockDim.x) {
    float o = 0.f;
    for (int w2 = 0; w2 < nw; ++w2) o += part[w2 * dh + d];
    out[(long long)t * D + h * dh + d] = o / sum;
  }
}
int launch_probe_attention(const float* q, const __nv_bfloat16* kv, float* out, int T, int S, int H, int dh, cudaStream_t s) {
  if (T <= 0) return 0;
  if (dh > 128 || dh % 2 != 0 || S <= 0 || H <= 0) return -2;
  const int threads = 256;
  const size_t smem = (size_t)(S + (threads / 32) * dh) * sizeof(float);
  if (smem > 48 * 1024) return -2;
  probe_attention_kernel<<<dim3(T, H), threads, smem, s>>>(q, kv, out, S, H, dh);
  return 0;
}

// out[row, :] = bf16(sum_s partial[s][row, :] + bias)   (generic split-K finish for small-M linear layers)
__global__ void splitk_finish_bf16_kernel(const float* __restrict__ partial, int n_planes, long long plane_stride,
                                          const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int N, int act) {
  const long long row = blockIdx.x;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float v = bias ? __ldg(bias + c) : 0.f;
    for (int p = 0; p < n_planes; ++p) v += __ldg(partial + p * plane_stride + row * N + c);
    if (act == 2) v = 0.5f * v * (1.0f + erff(v * 0.7071067811865476f));
    else if (act == 1) v = 0.5f * v * (1.0f + tanhf(0.7978845608028654f * (v + 0.044715f * v * v * v)));
    out[row * N + c] = __float2bfloat16_rn(v);
  }
}
int launch_splitk_finish_bf16(const float* partial, int n_planes, long long plane_stride, const float* bias,
                              __nv_bfloat16* out, long long rows, int N, int act, cudaStream_t s) {
  if (rows <= 0) return 0;
  splitk_finish_bf16_kernel<<<(unsigned)rows, 256, 0, s>>>(partial, n_planes, plane_stride, bias, out, N, act);
  return 0;
}

}  // namespace mmd
