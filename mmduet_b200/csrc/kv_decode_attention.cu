// KV-append attention in the DECODE regime: one or two new tokens of a stream (R = n_q x G <= 16 stacked query rows per KV
// head; greedy generation is n_q = 1) attend to a long paged context.  7 FLOP per KV byte: this is the one HBM-bound attention
// shape on the path (north star: ">= 70 % of HBM peak on KV-append attention"), so the kernel is built around the KV stream:
//   * a CTA owns one (stream, KV head, key range); the 64-token pages of its range (two contiguous 16 KB blocks each: K, V)
//     are fetched with 16-B vectorised, fully coalesced cp.async into a 4-deep shared-memory ring (128 KB in flight per SM);
//   * the four warps split every page's 64 keys (16 each): S = Q K^T and O += P V on mma.sync.m16n8k16 with the <= 16 query
//     rows as the M dimension (Q fragments live in registers for the whole kernel), fp32 online softmax with warp-shuffle row
//     reductions (two xor steps inside the quad);
//   * the warps' partial (m, l, O) are merged in shared memory and written as one split-KV partial; the existing combine
//     kernel (kv_attention.cu, up to 64 splits) finishes, so a single stream still fills all SMs.
// Replaces SDPA under Qwen2Attention for q_len = 1 (TF:models/qwen2/modeling_qwen2.py:187-246; the token loop of
// models/modeling_live.py:51-77).
#include "kernels.cuh"
#include "launch.cuh"

#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

namespace mmd {

namespace {

constexpr int KD_ROWS = 16;      // stacked query rows (token x group head), the MMA M dimension
constexpr int KD_BN = 64;        // keys per tile == tokens per KV page
constexpr int KD_DH = 128;
constexpr int KD_LDS = 136;      // padded smem row (bf16 elements): conflict-free ldmatrix
constexpr int KD_WARPS = 4, KD_THREADS = 32 * KD_WARPS;
constexpr int kd_smem(int stages) { return (KD_ROWS + 2 * stages * KD_BN) * KD_LDS * 2; }

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

template <int KD_STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(KD_THREADS, MIN_CTAS)
kv_decode_attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv_layer,
                           const int* __restrict__ stream_desc, const int* __restrict__ block_tables, float* __restrict__ o_part,
                           float* __restrict__ ml_part, int Hq, int Hkv, int n_splits, long long part_stride_rows, float scale_log2e) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + KD_ROWS * KD_LDS;
  __nv_bfloat16* sV = sK + KD_STAGES * KD_BN * KD_LDS;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int G = Hq / Hkv;
  const int kvh = blockIdx.x, sp = blockIdx.y, st = blockIdx.z;
  const int q_start = stream_desc[st * 4 + 0], n_q = stream_desc[st * 4 + 1], kv_len = stream_desc[st * 4 + 2];
  const int* table = block_tables + stream_desc[st * 4 + 3];
  const int R = n_q * G;                       // <= KD_ROWS (checked by the launcher's caller)
  const int past = kv_len - n_q;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (kv_len + KD_BN - 1) / KD_BN;
  const int per = (n_tiles + n_splits - 1) / n_splits;
  const int t_begin = sp * per, t_end = min(n_tiles, t_begin + per);

  constexpr int CH = KD_DH / 8;                // 16-B chunks per row
  for (int i = tid; i < KD_ROWS * CH; i += KD_THREADS) {
    const int r = i / CH, c = i % CH;
    const bool ok = r < R;
    const int tok = ok ? r / G : 0, g = ok ? r % G : 0;
    cp_async16((uint32_t)__cvta_generic_to_shared(sQ + r * KD_LDS + c * 8), q + ((long long)(q_start + tok) * Hq + kvh * G + g) * KD_DH + c * 8,
               ok ? 16 : 0);
  }
  auto load_kv = [&](int tile, int buf) {      // one page: two contiguous 16 KB blocks, 8 x 16 B per thread each
    const int page = table[tile];
    const __nv_bfloat16* gK = kv_layer + (((long long)page * 2 + 0) * Hkv + kvh) * KD_BN * KD_DH;
    const __nv_bfloat16* gV = kv_layer + (((long long)page * 2 + 1) * Hkv + kvh) * KD_BN * KD_DH;
    const int valid = kv_len - tile * KD_BN;   // rows >= valid are zero-filled (pool memory may hold anything)
#pragma unroll
    for (int j = 0; j < KD_BN * CH / KD_THREADS; ++j) {
      const int i = j * KD_THREADS + tid, r = i / CH, c = i % CH;
      const int nbytes = r < valid ? 16 : 0;
      cp_async16((uint32_t)__cvta_generic_to_shared(sK + (buf * KD_BN + r) * KD_LDS + c * 8), gK + r * KD_DH + c * 8, nbytes);
      cp_async16((uint32_t)__cvta_generic_to_shared(sV + (buf * KD_BN + r) * KD_LDS + c * 8), gV + r * KD_DH + c * 8, nbytes);
    }
  };
#pragma unroll
  for (int s = 0; s < KD_STAGES - 1; ++s) {    // prologue: STAGES-1 pages in flight (the first group also carries Q)
    if (t_begin + s < t_end) load_kv(t_begin + s, s);
    cp_async_commit();
  }

  constexpr int KSTEPS = KD_DH / 16, NT_O = KD_DH / 8;
  uint32_t qf[KSTEPS][4];
  bool q_loaded = false;
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int row0 = lane >> 2, row1 = row0 + 8;
  const int lim0 = past + min(row0, R - 1) / G, lim1 = past + min(row1, R - 1) / G;   // last key position each row may see
  const int kw0 = warp * 16;                   // this warp's 16 keys of every tile

  for (int tile = t_begin; tile < t_end; ++tile) {
    const int buf = (tile - t_begin) % KD_STAGES;
    {   // keep STAGES-1 pages in flight: the slot being refilled was consumed in the previous iteration (barrier below)
      const int nxt = tile + KD_STAGES - 1;
      if (nxt < t_end) load_kv(nxt, (nxt - t_begin) % KD_STAGES);
      cp_async_commit();
    }
    cp_async_wait<KD_STAGES - 1>();
    __syncthreads();
    if (!q_loaded) {
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int r = (lane & 7) + ((lane >> 3) & 1) * 8, c = ks * 16 + (lane >> 4) * 8;
        ldmatrix_x4((uint32_t)__cvta_generic_to_shared(sQ + r * KD_LDS + c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
      q_loaded = true;
    }
    const __nv_bfloat16* k_s = sK + buf * KD_BN * KD_LDS;
    const __nv_bfloat16* v_s = sV + buf * KD_BN * KD_LDS;
    float sc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      uint32_t b0, b1, b2, b3;
      const int kr = kw0 + (lane & 7) + (lane >> 4) * 8, kc = ks * 16 + ((lane >> 3) & 1) * 8;
      ldmatrix_x4((uint32_t)__cvta_generic_to_shared(k_s + kr * KD_LDS + kc), b0, b1, b2, b3);
      mma_bf16_16816(sc[0], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
      mma_bf16_16816(sc[1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
    }
    {   // causal / length mask (only the last tile or two can reach beyond a row's limit)
      const int kb = tile * KD_BN + kw0 + (lane & 3) * 2;
      if (kb + 15 > min(lim0, lim1)) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int key = kb + nt * 8;
          if (key > lim0) sc[nt][0] = -INFINITY;
          if (key + 1 > lim0) sc[nt][1] = -INFINITY;
          if (key > lim1) sc[nt][2] = -INFINITY;
          if (key + 1 > lim1) sc[nt][3] = -INFINITY;
        }
      }
    }
    float mx[2] = {fmaxf(fmaxf(sc[0][0], sc[0][1]), fmaxf(sc[1][0], sc[1][1])), fmaxf(fmaxf(sc[0][2], sc[0][3]), fmaxf(sc[1][2], sc[1][3]))};
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;   // fully masked so far: avoid (-inf) - (-inf)
      corr[r] = exp2f((m_run[r] - m_use) * scale_log2e);
      m_run[r] = m_new;
      msc[r] = m_use * scale_log2e;
      l_run[r] *= corr[r];
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      sc[nt][0] = exp2f(sc[nt][0] * scale_log2e - msc[0]);
      sc[nt][1] = exp2f(sc[nt][1] * scale_log2e - msc[0]);
      sc[nt][2] = exp2f(sc[nt][2] * scale_log2e - msc[1]);
      sc[nt][3] = exp2f(sc[nt][3] * scale_log2e - msc[1]);
      l_run[0] += sc[nt][0] + sc[nt][1];
      l_run[1] += sc[nt][2] + sc[nt][3];
    }
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    const uint32_t a0 = pack_bf16x2(sc[0][0], sc[0][1]), a1 = pack_bf16x2(sc[0][2], sc[0][3]);
    const uint32_t a2 = pack_bf16x2(sc[1][0], sc[1][1]), a3 = pack_bf16x2(sc[1][2], sc[1][3]);
#pragma unroll
    for (int np = 0; np < NT_O / 2; ++np) {
      uint32_t b0, b1, b2, b3;
      const int vr = kw0 + (lane & 7) + ((lane >> 3) & 1) * 8, vc = np * 16 + (lane >> 4) * 8;
      ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(v_s + vr * KD_LDS + vc), b0, b1, b2, b3);
      mma_bf16_16816(o[2 * np], a0, a1, a2, a3, b0, b1);
      mma_bf16_16816(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
    __syncthreads();       // every warp is done with this slot before the next iteration refills its predecessor
  }
  cp_async_wait<0>();
  __syncthreads();

  // merge the four warps' partial results (each saw a quarter of the keys) through shared memory, reusing the ring
  float* sO = reinterpret_cast<float*>(sK);                  // [warp][16 rows][128]
  float* sM = sO + KD_WARPS * KD_ROWS * KD_DH;               // [warp][16 rows][2]
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    const int row = r ? row1 : row0;
    float* op = sO + (warp * KD_ROWS + row) * KD_DH;
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt) *reinterpret_cast<float2*>(op + nt * 8 + (lane & 3) * 2) = make_float2(o[nt][2 * r], o[nt][2 * r + 1]);
    if ((lane & 3) == 0) {
      sM[(warp * KD_ROWS + row) * 2 + 0] = m_run[r];
      sM[(warp * KD_ROWS + row) * 2 + 1] = l_run[r];
    }
  }
  __syncthreads();
  {
    const int row = tid >> 3, c0 = (tid & 7) * 16;           // 128 threads = 16 rows x 8 column blocks of 16
    if (row < R) {
      float m = -INFINITY;
#pragma unroll
      for (int w = 0; w < KD_WARPS; ++w) m = fmaxf(m, sM[(w * KD_ROWS + row) * 2]);
      float wgt[KD_WARPS], l = 0.f;
#pragma unroll
      for (int w = 0; w < KD_WARPS; ++w) {
        const float mw = sM[(w * KD_ROWS + row) * 2];
        wgt[w] = (mw == -INFINITY) ? 0.f : exp2f((mw - m) * scale_log2e);
        l += wgt[w] * sM[(w * KD_ROWS + row) * 2 + 1];
      }
      const int tok = row / G, g = row % G;
      const long long grow = (long long)(q_start + tok) * Hq + kvh * G + g;
      float* op = o_part + ((long long)sp * part_stride_rows + grow) * KD_DH + c0;
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < KD_WARPS; ++w) {
          const float4 v = *reinterpret_cast<const float4*>(sO + (w * KD_ROWS + row) * KD_DH + c0 + c);
          acc.x += wgt[w] * v.x; acc.y += wgt[w] * v.y; acc.z += wgt[w] * v.z; acc.w += wgt[w] * v.w;
        }
        *reinterpret_cast<float4*>(op + c) = acc;
      }
      if ((tid & 7) == 0) {
        float* mp = ml_part + ((long long)sp * part_stride_rows + grow) * 2;
        mp[0] = (m == -INFINITY) ? -INFINITY : m * scale_log2e;
        mp[1] = l;
      }
    }
  }
}

}  // namespace

bool kv_decode_applicable(int max_rows) { return max_rows <= KD_ROWS; }

// shape of the launch: 0 = one CTA per SM with a 5-page ring, 1 = two CTAs per SM with 3-page rings (MMD_KD_VARIANT, measured)
static int kd_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MMD_KD_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

// as many splits as fill the GPU, at least one page each, at most what the combine kernel merges (64)
int kv_decode_pick_splits(int Hkv, int n_streams, int max_kv_len, int num_sms) {
  const int tiles = (max_kv_len + KD_BN - 1) / KD_BN;
  int s = (kd_variant() == 1 ? 2 : 1) * num_sms / (Hkv * n_streams > 0 ? Hkv * n_streams : 1);
  if (s > tiles) s = tiles;
  if (s > 64) s = 64;
  return s < 1 ? 1 : s;
}

int launch_kv_decode_attention(const __nv_bfloat16* q, const __nv_bfloat16* kv_layer, const int* stream_desc, const int* block_tables,
                               int n_streams, int total_q, float* o_part, float* ml_part, int Hq, int Hkv, int n_splits, cudaStream_t s) {
  static PerDeviceFlag attr;
  if (!attr.cur()) {
    if (cudaFuncSetAttribute(kv_decode_attention_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kd_smem(5)) != cudaSuccess) return -4;
    if (cudaFuncSetAttribute(kv_decode_attention_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kd_smem(3)) != cudaSuccess) return -4;
    attr.cur() = true;
  }
  const float scale_log2e = (1.0f / sqrtf((float)KD_DH)) * 1.4426950408889634f;
  if (kd_variant() == 1)
    launch_k(kv_decode_attention_kernel<3, 2>, dim3(Hkv, n_splits, n_streams), dim3(KD_THREADS), kd_smem(3), s, q, kv_layer, stream_desc,
             block_tables, o_part, ml_part, Hq, Hkv, n_splits, (long long)total_q * Hq, scale_log2e);
  else
    launch_k(kv_decode_attention_kernel<5, 1>, dim3(Hkv, n_splits, n_streams), dim3(KD_THREADS), kd_smem(5), s, q, kv_layer, stream_desc,
             block_tables, o_part, ml_part, Hq, Hkv, n_splits, (long long)total_q * Hq, scale_log2e);
  return 0;
}

}  // namespace mmd
