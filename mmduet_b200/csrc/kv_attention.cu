// KV-append ("chunked prefill") attention of the Qwen2 decoder over a paged KV pool.
// For every stream the step's M new tokens (already appended to the pool by the QKV-finish kernel) attend to all
// L = past + M keys with a bottom-right causal mask.  The G = Hq/Hkv query heads of a KV head are stacked into one
// row block (GQA: 7 x 49 = 343 rows for a frame step), so each K/V page is streamed once per 128-row tile:
//   * K/V pages (64 tokens x 128 dims, contiguous 16 KB) are read with 16-B cp.async, double-buffered;
//   * QK^T and PV run on mma.sync.m16n8k16 (bf16 in, fp32 accumulate) — at 343 FLOP/B this kernel is bound by the
//     tensor pipe, not by HBM, CUDA-core FMAs would reach only ~3% of HBM bandwidth;
//   * online softmax in fp32 registers with warp-shuffle row reductions;
//   * split-KV across CTAs (flash-decoding) with a small combine kernel, so one stream fills the GPU.
// Replaces SDPA / flash-attn-2 under Qwen2Attention (TF:models/qwen2/modeling_qwen2.py:187-246,
// TF:integrations/sdpa_attention.py:41-104) and the O(L) torch.cat of DynamicCache.update (TF:cache_utils.py:119-120).
#include "kernels.cuh"
#include "launch.cuh"

#include <cuda_bf16.h>
#include <math.h>

namespace mmd {

namespace {

constexpr int KA_BM = 128;      // query rows (token x group-head) per CTA
constexpr int KA_BN = 64;       // keys per tile == tokens per KV page
constexpr int KA_DH = 128;
constexpr int KA_LDS = 136;     // padded smem row (bf16 elements): conflict-free ldmatrix
constexpr int KA_THREADS = 256;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// stream descriptor: {q_start (first row of this stream in the packed token buffer), n_q, kv_len (past + n_q),
//                     table_off (offset into block_tables)}
__global__ void __launch_bounds__(KA_THREADS, 2)
kv_attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv_layer,
                    const int* __restrict__ stream_desc, const int* __restrict__ block_tables,
                    float* __restrict__ o_part, float* __restrict__ ml_part, int Hq, int Hkv, int n_splits,
                    long long part_stride_rows, float scale_log2e) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + KA_BM * KA_LDS;
  __nv_bfloat16* sV = sK + 2 * KA_BN * KA_LDS;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int G = Hq / Hkv;
  const int kvh = blockIdx.x % Hkv, qt = blockIdx.x / Hkv;
  const int sp = blockIdx.y, st = blockIdx.z;
  const int q_start = stream_desc[st * 4 + 0], n_q = stream_desc[st * 4 + 1];
  const int kv_len = stream_desc[st * 4 + 2];
  const int* table = block_tables + stream_desc[st * 4 + 3];
  const int R = n_q * G;
  const int r_base = qt * KA_BM;
  if (r_base >= R) return;
  const int past = kv_len - n_q;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // key tiles this CTA covers: split `sp` of the tiles visible to the last row of this q tile
  const int last_row = min(R, r_base + KA_BM) - 1;
  const int max_pos = past + last_row / G;
  const int n_tiles_vis = min((kv_len + KA_BN - 1) / KA_BN, max_pos / KA_BN + 1);
  const int per = (n_tiles_vis + n_splits - 1) / n_splits;
  const int t_begin = sp * per, t_end = min(n_tiles_vis, t_begin + per);
  const int min_pos = past + r_base / G;

  constexpr int CH = KA_DH / 8;
  for (int i = tid; i < KA_BM * CH; i += KA_THREADS) {
    const int r = i / CH, c = i % CH;
    const int rr = r_base + r;
    const bool ok = rr < R;
    const int tok = ok ? rr / G : 0, g = ok ? rr % G : 0;
    const __nv_bfloat16* src = q + ((long long)(q_start + tok) * Hq + kvh * G + g) * KA_DH + c * 8;
    cp_async16((uint32_t)__cvta_generic_to_shared(sQ + r * KA_LDS + c * 8), src, ok ? 16 : 0);
  }
  auto load_kv = [&](int tile, int buf) {
    const int page = table[tile];
    const __nv_bfloat16* gK = kv_layer + (((long long)page * 2 + 0) * Hkv + kvh) * KA_BN * KA_DH;
    const __nv_bfloat16* gV = kv_layer + (((long long)page * 2 + 1) * Hkv + kvh) * KA_BN * KA_DH;
    const int valid = kv_len - tile * KA_BN;  // rows >= valid are zero-filled (pool memory may hold anything)
    for (int i = tid; i < KA_BN * CH; i += KA_THREADS) {
      const int r = i / CH, c = i % CH;
      const int nbytes = r < valid ? 16 : 0;
      cp_async16((uint32_t)__cvta_generic_to_shared(sK + (buf * KA_BN + r) * KA_LDS + c * 8), gK + r * KA_DH + c * 8, nbytes);
      cp_async16((uint32_t)__cvta_generic_to_shared(sV + (buf * KA_BN + r) * KA_LDS + c * 8), gV + r * KA_DH + c * 8, nbytes);
    }
  };
  if (t_begin < t_end) load_kv(t_begin, 0);
  cp_async_commit();

  constexpr int KSTEPS = KA_DH / 16, NT_S = KA_BN / 8, NT_O = KA_DH / 8;
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  // causal limit of this thread's two rows (row = r_base + warp*16 + lane/4 (+8)); rows >= R get limit of row R-1
  const int row0 = r_base + warp * 16 + (lane >> 2), row1 = row0 + 8;
  const int lim0 = past + min(row0, R - 1) / G, lim1 = past + min(row1, R - 1) / G;

  for (int tile = t_begin; tile < t_end; ++tile) {
    const int buf = (tile - t_begin) & 1;
    if (tile + 1 < t_end) load_kv(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const __nv_bfloat16* k_s = sK + buf * KA_BN * KA_LDS;
    const __nv_bfloat16* v_s = sV + buf * KA_BN * KA_LDS;
    float sc[NT_S][4];
#pragma unroll
    for (int i = 0; i < NT_S; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      uint32_t a0, a1, a2, a3;
      {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = ks * 16 + (lane >> 4) * 8;
        ldmatrix_x4((uint32_t)__cvta_generic_to_shared(sQ + r * KA_LDS + c), a0, a1, a2, a3);
      }
#pragma unroll
      for (int np = 0; np < NT_S / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int kr = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int kc = ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4((uint32_t)__cvta_generic_to_shared(k_s + kr * KA_LDS + kc), b0, b1, b2, b3);
        mma_bf16_16816(sc[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(sc[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // causal / length mask: only tiles that reach beyond the first row's position need it (CTA-uniform test)
    const int k0 = tile * KA_BN;
    if (k0 + KA_BN - 1 > min_pos) {
      const int kb = k0 + (lane & 3) * 2;
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt) {
        const int key = kb + nt * 8;
        if (key > lim0) sc[nt][0] = -INFINITY;
        if (key + 1 > lim0) sc[nt][1] = -INFINITY;
        if (key > lim1) sc[nt][2] = -INFINITY;
        if (key + 1 > lim1) sc[nt][3] = -INFINITY;
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(sc[nt][0], sc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(sc[nt][2], sc[nt][3]));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked so far: avoid (-inf) - (-inf)
      corr[r] = exp2f((m_run[r] - m_use) * scale_log2e);
      m_run[r] = m_new;
      msc[r] = m_use * scale_log2e;
      l_run[r] *= corr[r];
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
      sc[nt][0] = exp2f(sc[nt][0] * scale_log2e - msc[0]);
      sc[nt][1] = exp2f(sc[nt][1] * scale_log2e - msc[0]);
      sc[nt][2] = exp2f(sc[nt][2] * scale_log2e - msc[1]);
      sc[nt][3] = exp2f(sc[nt][3] * scale_log2e - msc[1]);
      rs[0] += sc[nt][0] + sc[nt][1];
      rs[1] += sc[nt][2] + sc[nt][3];
    }
    l_run[0] += rs[0];
    l_run[1] += rs[1];
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
#pragma unroll
    for (int kk = 0; kk < KA_BN / 16; ++kk) {
      const uint32_t a0 = pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]);
      const uint32_t a1 = pack_bf16x2(sc[2 * kk][2], sc[2 * kk][3]);
      const uint32_t a2 = pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
      const uint32_t a3 = pack_bf16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < NT_O / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int vr = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int vc = np * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(v_s + vr * KA_LDS + vc), b0, b1, b2, b3);
        mma_bf16_16816(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // write the un-normalised partial output and (m, l) of this split
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = r ? row1 : row0;
    if (row < R) {
      const int tok = row / G, g = row % G;
      const long long grow = (long long)(q_start + tok) * Hq + kvh * G + g;
      float* op = o_part + ((long long)sp * part_stride_rows + grow) * KA_DH;
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt) {
        const int c = nt * 8 + (lane & 3) * 2;
        *reinterpret_cast<float2*>(op + c) = make_float2(o[nt][2 * r], o[nt][2 * r + 1]);
      }
      if ((lane & 3) == 0) {
        float* mp = ml_part + ((long long)sp * part_stride_rows + grow) * 2;
        mp[0] = (m_run[r] == -INFINITY) ? -INFINITY : m_run[r] * scale_log2e;
        mp[1] = l_run[r];
      }
    }
  }
}

// out[tok, head*128 + d] = sum_s w_s O_s / sum_s w_s l_s,  w_s = 2^(m_s - max m)
__global__ void kv_attention_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part,
                                            __nv_bfloat16* __restrict__ out, int n_splits, long long part_stride_rows) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // one warp per (token, head) row: lane l owns dims 4l..4l+3 (one 16-B load per split, one 8-B store).  The (m, l)
  // pairs of the <= 32 splits are read by one lane each and combined with shuffles, so the weights are known before
  // the O loads are issued and those can all be in flight together.
  const long long grow = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (grow >= part_stride_rows) return;
  const int lane = threadIdx.x & 31;
  // lane l owns the (m, l) pairs of splits l and l + 32 (up to 64 splits: the decode kernel spreads one stream over all SMs)
  float2 ml = make_float2(-INFINITY, 0.f), ml2 = make_float2(-INFINITY, 0.f);
  if (lane < n_splits) ml = __ldg(reinterpret_cast<const float2*>(ml_part + ((long long)lane * part_stride_rows + grow) * 2));
  if (lane + 32 < n_splits) ml2 = __ldg(reinterpret_cast<const float2*>(ml_part + ((long long)(lane + 32) * part_stride_rows + grow) * 2));
  float mmax = fmaxf(ml.x, ml2.x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mmax = fmaxf(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
  const float w_mine = (ml.x == -INFINITY) ? 0.f : exp2f(ml.x - mmax);
  const float w_mine2 = (ml2.x == -INFINITY) ? 0.f : exp2f(ml2.x - mmax);
  float lsum = w_mine * ml.y + w_mine2 * ml2.y;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int s = 0; s < n_splits; ++s) {
    const float w = s < 32 ? __shfl_sync(0xffffffffu, w_mine, s) : __shfl_sync(0xffffffffu, w_mine2, s - 32);
    const float4 o = __ldg(reinterpret_cast<const float4*>(o_part + ((long long)s * part_stride_rows + grow) * KA_DH) + lane);
    acc.x += w * o.x; acc.y += w * o.y; acc.z += w * o.z; acc.w += w * o.w;
  }
  const float inv = 1.f / lsum;
  __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x * inv, acc.y * inv), hi = __floats2bfloat162_rn(acc.z * inv, acc.w * inv);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(out + grow * KA_DH + lane * 4) = pk;
}

// The same merge for FEW rows and MANY splits (token-by-token decode: 28 rows x up to 64 splits): one block per row, the
// eight warps take the splits round-robin so that every partial row is in flight at once (a single warp walking 37 splits
// in order costs ~7 us of dependent L2 latency), then a shared-memory reduction.  Same arithmetic order per warp as above
// is not required: fp32 sums of <= 64 non-negative-weighted terms, the result is rounded to bf16.
__global__ void kv_attention_combine_wide_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part,
                                                 __nv_bfloat16* __restrict__ out, int n_splits, long long part_stride_rows) {
  __shared__ float sw[64];
  __shared__ float4 sacc[8][32];
  __shared__ float s_l;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long grow = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    float2 ml = make_float2(-INFINITY, 0.f), ml2 = make_float2(-INFINITY, 0.f);
    if (lane < n_splits) ml = __ldg(reinterpret_cast<const float2*>(ml_part + ((long long)lane * part_stride_rows + grow) * 2));
    if (lane + 32 < n_splits) ml2 = __ldg(reinterpret_cast<const float2*>(ml_part + ((long long)(lane + 32) * part_stride_rows + grow) * 2));
    float mmax = fmaxf(ml.x, ml2.x);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mmax = fmaxf(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
    const float w1 = (ml.x == -INFINITY) ? 0.f : exp2f(ml.x - mmax), w2 = (ml2.x == -INFINITY) ? 0.f : exp2f(ml2.x - mmax);
    float lsum = w1 * ml.y + w2 * ml2.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    sw[lane] = w1;
    sw[lane + 32] = w2;
    if (lane == 0) s_l = lsum;
  }
  // the O rows do not depend on the weights: issue this warp's loads (<= 8 splits) before waiting for them
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = warp + 8 * i;
    v[i] = s < n_splits ? __ldg(reinterpret_cast<const float4*>(o_part + ((long long)s * part_stride_rows + grow) * KA_DH) + lane)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = warp + 8 * i;
    const float w = s < n_splits ? sw[s] : 0.f;
    acc.x += w * v[i].x; acc.y += w * v[i].y; acc.z += w * v[i].z; acc.w += w * v[i].w;
  }
  sacc[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 a = sacc[w][lane];
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    const float inv = 1.f / s_l;
    __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x * inv, acc.y * inv), hi = __floats2bfloat162_rn(acc.z * inv, acc.w * inv);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + grow * KA_DH + lane * 4) = pk;
  }
}

}  // namespace

// 0 = mma.sync kernels, 1 = tcgen05/TMEM kernels (attn_tcgen05.cu), 2 = auto: tcgen05 for the ViT and for decoder steps
// with >= 1024 stacked query rows or >= 2k context (measured faster from there on: tools/long_stream.py), mma.sync for
// short single-frame steps.
int g_attention_impl = 2;

static bool kv_use_tc(int max_rows, int max_kv_len) {
  if (g_attention_impl == 2) return max_rows >= 1024 || max_kv_len >= 2048;
  return g_attention_impl == 1;
}

int launch_kv_attention_tc_main(const __nv_bfloat16* q, const __nv_bfloat16* kv_layer, const int* stream_desc, const int* block_tables,
                                int n_streams, int max_n_q, int total_q, float* o_part, float* ml_part, int Hq, int Hkv, int n_splits,
                                cudaStream_t s);

bool kv_decode_applicable(int max_rows);
int kv_decode_pick_splits(int Hkv, int n_streams, int max_kv_len, int num_sms);
int launch_kv_decode_attention(const __nv_bfloat16* q, const __nv_bfloat16* kv_layer, const int* stream_desc, const int* block_tables,
                               int n_streams, int total_q, float* o_part, float* ml_part, int Hq, int Hkv, int n_splits, cudaStream_t s);
// 1 (default): passes with <= 16 stacked query rows per KV head (token-by-token generation) take the HBM-streaming decode kernel
int g_kv_decode = 1;

int kv_attention_pick_splits(int max_rows, int Hkv, int n_streams, int max_kv_len, int num_sms) {
  if (g_kv_decode && kv_decode_applicable(max_rows)) return kv_decode_pick_splits(Hkv, n_streams, max_kv_len, num_sms);
  // mma.sync kernel: 128 query rows per CTA, two CTAs per SM; tcgen05 kernel: 256 rows per CTA, one CTA per SM
  const bool tc = kv_use_tc(max_rows, max_kv_len);
  const int rows_per_cta = tc ? 2 * KA_BM : KA_BM;
  const int q_tiles = (max_rows + rows_per_cta - 1) / rows_per_cta;
  const int base = q_tiles * Hkv * n_streams;
  const int kv_tiles = (max_kv_len + KA_BN - 1) / KA_BN;
  // Pick the split count that minimises (waves of CTAs) x (key tiles per CTA + fixed per-CTA cost): a count that spills
  // a few CTAs into a second wave doubles the kernel time, so "just fill the SMs" is the wrong rule.  Fixed cost in
  // key-tile units (measured, tools/trace_attn.py): prologue + epilogue of the tcgen05 kernel ~ 8 tiles, mma.sync ~ 2.
  const int slots = (tc ? 1 : 2) * num_sms;
  const double fixed = tc ? 8.0 : 2.0;
  int max_by_work = (kv_tiles + 3) / 4;  // at least ~4 key tiles (256 keys) per split
  if (max_by_work > 32) max_by_work = 32;
  if (max_by_work < 1) max_by_work = 1;
  int splits = 1;
  double best = 1e30;
  for (int sp = 1; sp <= max_by_work; ++sp) {
    const int waves = (base * sp + slots - 1) / slots;
    const int per = (kv_tiles + sp - 1) / sp;
    const double cost = waves * (per + fixed) + 0.1 * sp;    // + the combine kernel reading sp partial planes
    if (cost < best - 1e-9) { best = cost; splits = sp; }
  }
  return splits;
}

int launch_kv_attention(const __nv_bfloat16* q, const __nv_bfloat16* kv_layer, const int* stream_desc, const int* block_tables,
                        int n_streams, int max_n_q, int total_q, int max_kv_len, float* o_part, float* ml_part, __nv_bfloat16* out,
                        int Hq, int Hkv, int dh, int page_tokens, int n_splits, cudaStream_t s) {
  if (n_streams <= 0 || total_q <= 0) return 0;
  if (dh != KA_DH || page_tokens != KA_BN || Hq % Hkv != 0 || n_splits < 1 || n_splits > 64) return -2;   // combine: two splits per lane
  constexpr int SMEM = (KA_BM + 4 * KA_BN) * KA_LDS * 2;
  static PerDeviceFlag attr;
  if (!attr.cur()) {
    if (cudaFuncSetAttribute(kv_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return -4;
    attr.cur() = true;
  }
  const int G = Hq / Hkv;
  const int q_tiles = (max_n_q * G + KA_BM - 1) / KA_BM;
  dim3 grid(q_tiles * Hkv, n_splits, n_streams);
  const float scale_log2e = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  const long long part_rows = (long long)total_q * Hq;
  if (g_kv_decode && kv_decode_applicable(max_n_q * G)) {
    const int rc = launch_kv_decode_attention(q, kv_layer, stream_desc, block_tables, n_streams, total_q, o_part, ml_part, Hq, Hkv, n_splits, s);
    if (rc != 0) return rc;
  } else if (n_splits > 32) {
    return -2;
  } else if (kv_use_tc(max_n_q * G, max_kv_len)) {
    const int rc = launch_kv_attention_tc_main(q, kv_layer, stream_desc, block_tables, n_streams, max_n_q, total_q, o_part, ml_part, Hq,
                                               Hkv, n_splits, s);
    if (rc != 0) return rc;
  } else {
    launch_k(kv_attention_kernel, grid, dim3(KA_THREADS), SMEM, s, q, kv_layer, stream_desc, block_tables, o_part, ml_part, Hq, Hkv,
             n_splits, part_rows, scale_log2e);
  }
  if (part_rows <= 512 && n_splits > 4)   // few rows, many splits: one block per row, all partial rows in flight
    launch_k(kv_attention_combine_wide_kernel, dim3((unsigned)part_rows), dim3(256), 0, s, o_part, ml_part, out, n_splits, part_rows);
  else
    launch_k(kv_attention_combine_kernel, dim3((unsigned)((part_rows + 7) / 8)), dim3(256), 0, s, o_part, ml_part, out, n_splits, part_rows);
  return 0;
}

}  // namespace mmd
