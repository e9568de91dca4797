// Fused non-causal attention over the 729 patch tokens of each frame (SigLIP, 16 heads x head_dim 72):
// softmax(Q K^T / sqrt(dh)) V without materialising the [T,16,729,729] score tensor that the reference's eager
// attention writes (TF:models/siglip/modeling_siglip.py:252-330).  Flash-style: 128-query tile per CTA, 64-key tiles
// double-buffered with cp.async, QK^T and PV on mma.sync.m16n8k16 (bf16 in, fp32 accumulate), online softmax in fp32
// registers with warp-shuffle row reductions.  head_dim 72 is padded to 80 in shared memory (zero columns).
// TODO(perf): tcgen05/TMEM variant; this kernel is ~10% of the ViT FLOPs.
#include "kernels.cuh"
#include "launch.cuh"

#include <cuda_bf16.h>
#include <math.h>

namespace mmd {

namespace {

constexpr int VA_BM = 128;      // queries per CTA (8 warps x 16 rows)
constexpr int VA_BN = 64;       // keys per tile
constexpr int VA_THREADS = 256;

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

// SPLIT: the output row is [hi | lo] with hi = bf16(o) and lo = bf16(o - hi), so that the out-projection GEMM (run with
// K doubled against [W | W]) consumes the attention output at ~16 mantissa bits.  Rounding o to a single bf16 is the
// largest single contributor to the tower's deviation from the fp32 oracle (tools/noise_floor.py).
template <int DH, int DPAD, int LDS, bool SPLIT>
__global__ void __launch_bounds__(VA_THREADS, 2)
vit_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int S, int H, float scale_log2e) {
  // qkv: [T*S, 3*H*DH] rows = tokens, columns = [q | k | v], each [H, DH].  out: [T*S, H*DH] (or [T*S, 2*H*DH] if SPLIT).
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + VA_BM * LDS;
  __nv_bfloat16* sV = sK + 2 * VA_BN * LDS;
  const int q_tile = blockIdx.x, h = blockIdx.y, t = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row_stride = 3 * H * DH;
  const __nv_bfloat16* base = qkv + (long long)t * S * row_stride;
  const __nv_bfloat16* gQ = base + h * DH;
  const __nv_bfloat16* gK = base + (H + h) * DH;
  const __nv_bfloat16* gV = base + (2 * H + h) * DH;
  constexpr int CH = DH / 8;  // 16-B chunks per row

  // zero the pad columns [DH, DPAD) once; cp.async never writes them
  for (int i = tid; i < (VA_BM + 4 * VA_BN) * (DPAD - DH); i += VA_THREADS) {
    const int r = i / (DPAD - DH), c = DH + i % (DPAD - DH);
    sQ[r * LDS + c] = __float2bfloat16_rn(0.f);  // sQ, sK, sV are contiguous with the same row stride
  }

  const int q0 = q_tile * VA_BM;
  for (int i = tid; i < VA_BM * CH; i += VA_THREADS) {
    const int r = i / CH, c = i % CH;
    const int qr = q0 + r;
    const bool ok = qr < S;
    cp_async16((uint32_t)__cvta_generic_to_shared(sQ + r * LDS + c * 8), gQ + (long long)(ok ? qr : 0) * row_stride + c * 8, ok ? 16 : 0);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * VA_BN;
    for (int i = tid; i < VA_BN * CH; i += VA_THREADS) {
      const int r = i / CH, c = i % CH;
      const int kr = k0 + r;
      const bool ok = kr < S;
      const long long off = (long long)(ok ? kr : 0) * row_stride + c * 8;
      cp_async16((uint32_t)__cvta_generic_to_shared(sK + (buf * VA_BN + r) * LDS + c * 8), gK + off, ok ? 16 : 0);
      cp_async16((uint32_t)__cvta_generic_to_shared(sV + (buf * VA_BN + r) * LDS + c * 8), gV + off, ok ? 16 : 0);
    }
  };
  const int n_tiles = (S + VA_BN - 1) / VA_BN;
  load_kv(0, 0);
  cp_async_commit();

  constexpr int KSTEPS = DPAD / 16;   // QK^T k-steps
  constexpr int NT_S = VA_BN / 8;     // score n-tiles per key tile
  constexpr int NT_O = DPAD / 8;      // output n-tiles
  uint32_t qf[KSTEPS][4];
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int tile = 0; tile < n_tiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < n_tiles) load_kv(tile + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (tile == 0) {
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = ks * 16 + (lane >> 4) * 8;
        ldmatrix_x4((uint32_t)__cvta_generic_to_shared(sQ + r * LDS + c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const __nv_bfloat16* k_s = sK + buf * VA_BN * LDS;
    const __nv_bfloat16* v_s = sV + buf * VA_BN * LDS;
    float sc[NT_S][4];
#pragma unroll
    for (int i = 0; i < NT_S; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
      for (int np = 0; np < NT_S / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int kr = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int kc = ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4((uint32_t)__cvta_generic_to_shared(k_s + kr * LDS + kc), b0, b1, b2, b3);
        mma_bf16_16816(sc[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_bf16_16816(sc[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }
    // mask keys beyond S (last tile only) and update the running max / sum per query row
    const int kbase = tile * VA_BN + (lane & 3) * 2;
    if (tile == n_tiles - 1) {
#pragma unroll
      for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (kbase + nt * 8 + (e & 1) >= S) sc[nt][e] = -INFINITY;
        }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(sc[nt][0], sc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(sc[nt][2], sc[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], mx[r]);  // finite: every tile has at least one valid key
      corr[r] = exp2f((m_run[r] - m_new) * scale_log2e);
      m_run[r] = m_new;
      msc[r] = m_new * scale_log2e;
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
    // O += P V
#pragma unroll
    for (int kk = 0; kk < VA_BN / 16; ++kk) {
      const uint32_t a0 = pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]);
      const uint32_t a1 = pack_bf16x2(sc[2 * kk][2], sc[2 * kk][3]);
      const uint32_t a2 = pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
      const uint32_t a3 = pack_bf16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < NT_O / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const int vr = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int vc = np * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans((uint32_t)__cvta_generic_to_shared(v_s + vr * LDS + vc), b0, b1, b2, b3);
        mma_bf16_16816(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __syncthreads();  // all warps done with buf before it is refilled by the next iteration's prefetch
  }

  // finalise: divide by the row sums (quad-reduced) and store bf16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  const int out_stride = (SPLIT ? 2 : 1) * H * DH;
  __nv_bfloat16* ob = out + (long long)t * S * out_stride + h * DH;
  auto store2 = [&](int r, int c, float a, float b) {
    __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
    *reinterpret_cast<__nv_bfloat162*>(ob + (long long)r * out_stride + c) = hi;
    if (SPLIT) {
      const float2 hf = __bfloat1622float2(hi);
      *reinterpret_cast<__nv_bfloat162*>(ob + (long long)r * out_stride + H * DH + c) = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    }
  };
#pragma unroll
  for (int nt = 0; nt < NT_O; ++nt) {
    const int c = nt * 8 + (lane & 3) * 2;
    if (c < DH) {
      if (r0 < S) store2(r0, c, o[nt][0] * inv0, o[nt][1] * inv0);
      if (r1 < S) store2(r1, c, o[nt][2] * inv1, o[nt][3] * inv1);
    }
  }
}

}  // namespace

extern int g_attention_impl;
int launch_vit_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int T, int S, int H, int dh, int split_hi_lo, cudaStream_t s);

int launch_vit_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int T, int S, int H, int dh, int split_hi_lo, cudaStream_t s) {
  if (T <= 0) return 0;
  if (g_attention_impl != 0) return launch_vit_attention_tc(qkv, out, T, S, H, dh, split_hi_lo, s);
  if (dh != 72) return -2;  // SigLIP-so400m head_dim; other sizes need another instantiation
  constexpr int DH = 72, DPAD = 80, LDS = 88;
  constexpr int SMEM = (VA_BM + 4 * VA_BN) * LDS * 2;
  static PerDeviceFlag attr;
  if (!attr.cur()) {
    if (cudaFuncSetAttribute(vit_attention_kernel<DH, DPAD, LDS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return -4;
    if (cudaFuncSetAttribute(vit_attention_kernel<DH, DPAD, LDS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return -4;
    attr.cur() = true;
  }
  dim3 grid((S + VA_BM - 1) / VA_BM, H, T);
  const float scale_log2e = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  if (split_hi_lo) vit_attention_kernel<DH, DPAD, LDS, true><<<grid, VA_THREADS, SMEM, s>>>(qkv, out, S, H, scale_log2e);
  else vit_attention_kernel<DH, DPAD, LDS, false><<<grid, VA_THREADS, SMEM, s>>>(qkv, out, S, H, scale_log2e);
  return 0;
}

}  // namespace mmd
