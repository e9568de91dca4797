// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[lane, col] = sum_k X[lane, k] * Y[col, k]
//   * operands staged by TMA (cp.async.bulk.tensor, 128-B swizzle) into a multi-stage shared-memory ring,
//   * products issued by ONE thread with tcgen05.mma (UMMA 128 x BN x 16, kind::f16, fp32 accumulate),
//   * accumulators live in TMEM (two buffers, so the epilogue of tile i overlaps the main loop of tile i+1),
//   * epilogue warps read TMEM with tcgen05.ld and apply the fused epilogue (bias / GELU / residual / SwiGLU).
// Both operands are K-major ("TN" GEMM), which is exactly nn.Linear: activations [M,K] and weights [N,K].
//
// Replaces the cuBLAS calls underneath the reference's nn.Linear layers:
//   SigLIP q/k/v/out/fc1/fc2, patch-embed-as-GEMM, mm_projector (video_head_live_llava_qwen.py:90-98),
//   Qwen2 q/k/v/o/gate/up/down (video_head_live_llava_qwen.py:141-150).
#include "gemm.cuh"
#include "ptx.cuh"
#include "launch.cuh"

#include <mutex>
#include <string>
#include <unordered_map>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mmd {

static thread_local std::string g_gemm_err;
const char* gemm_last_error() { return g_gemm_err.c_str(); }

constexpr int BM = 128;        // UMMA M: TMEM lanes
constexpr int BK = 64;         // bf16 elements per k-block = one 128-B swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 8;   // two warps per TMEM lane quarter, each draining half of the accumulator columns
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;  // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2..9 epilogue
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int kMaxDevices = 16;     // per-device "attribute set" flags (power of two)

struct GemmKernelParams {
  int x_rows, y_rows;
  int x_tiles, y_tiles;
  int k_splits, kb_per_split, kb_total;
  const float* bias;
  void* out;
  long long ldo;
  long long split_stride;
  int pdl_prefetch_x;   // X operand (weights, swap-AB) may be loaded before griddepcontrol.wait
  int x_blocked;        // X/X2 tensor maps are 4-D over the tile-blocked weight layout
  int y_lo_off;         // HILO kernels: column offset (elements) of the lo half of the [hi | lo] Y operand (= K)
  int out_hilo;         // EPI_T_SWIGLU: write bf16 hi at [tok, n] and the rounding remainder lo at [tok, x_rows + n]
  const float* row_w;   // EPI_BF16_HILO_POOL: pooling weight of every X row
  int pool_group;       // EPI_BF16_HILO_POOL: X rows per pooled output row (4 or 16)
  int row_w_period;     // > 0: row_w is indexed by row % row_w_period
  float* raw_out;       // SwiGLU epilogues: raw fp32 (gate, up) pairs of the tokens >= raw_from (precise rows' hi / lo copies)
  int raw_from;
  long long ld_raw;
};

// HILO: the Y operand (activations, swap-AB) is a bf16 hi+lo pair [hi | lo] (2K wide): every stage carries the hi and the lo
// k-block of Y next to ONE weight tile and the MMA thread issues both products into the same accumulator, so the weights
// are streamed (HBM, L2 and shared memory) exactly once while the activation rounding error drops from 2^-9 to 2^-17.
template <int BN, bool DUAL, bool HILO = false>
struct GemmCfg {
  static constexpr int X_BYTES = BM * BK * 2;                      // 16 KB
  static constexpr int Y_HALF = BN * BK * 2;
  static constexpr int Y_BYTES = Y_HALF * (HILO ? 2 : 1);
  static constexpr int STAGE_BYTES = X_BYTES * (DUAL ? 2 : 1) + Y_BYTES;
  static constexpr int ACC_COLS = BN * (DUAL ? 2 : 1);
  static constexpr int TMEM_USED = 2 * ACC_COLS;                   // double-buffered accumulator
  static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int STAGES_RAW = (SMEM_BUDGET - 1024 /*align slack*/ - BAR_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
  static_assert(TMEM_USED <= 512, "TMEM overflow");
  static_assert(STAGES >= 3, "pipeline too shallow");
};

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh( sqrt(2/pi) (x + 0.044715 x^3) ))  -- SigLIP "gelu_pytorch_tanh"
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float gelu_erf_f(float x) {  // nn.GELU() default used by mm_projector (mlp2x_gelu)
  return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
}
template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if constexpr (ACT == ACT_GELU_TANH) return gelu_tanh_f(x);
  if constexpr (ACT == ACT_GELU_ERF) return gelu_erf_f(x);
  return x;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Normal-orientation epilogue of one accumulator tile: TMEM lane = output row `xi`, TMEM columns [c_begin, c_end) =
// output columns y0 + c.  Shared by the 1-CTA and the 2-CTA (cta_group::2) kernels.
template <int EPI, int ACT>
__device__ __forceinline__ void epilogue_normal(const GemmKernelParams& p, uint32_t t_row, int xi, int y0, int c_begin, int c_end) {
    // normal orientation: lane = output row m, columns = n (contiguous in memory)
    const bool row_ok = xi < p.x_rows;
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      if (y0 + c0 >= p.y_rows) break;  // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c0, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = y0 + c0 + g * 8;
          if (n < p.y_rows) {  // y_rows % 8 == 0 is enforced by the launcher
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
            if (p.bias != nullptr) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if constexpr (EPI == EPI_BF16 || EPI == EPI_BF16_HILO) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = apply_act<ACT>(f[j]);
              uint4 o;
              o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
              o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)xi * p.ldo + n;
              *reinterpret_cast<uint4*>(dst) = o;
              if constexpr (EPI == EPI_BF16_HILO) {
                const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&o);
                uint4 lo;
                uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 hf = __bfloat1622float2(hp[j]);
                  lw[j] = pack_bf16(f[2 * j] - hf.x, f[2 * j + 1] - hf.y);
                }
                *reinterpret_cast<uint4*>(dst + p.y_rows) = lo;
              }
            } else {
              float* dst = reinterpret_cast<float*>(p.out) + (long long)xi * p.ldo + n;
              float4 r0, r1;
              if constexpr (EPI == EPI_RESID_F32) {
                r0 = *reinterpret_cast<const float4*>(dst);
                r1 = *reinterpret_cast<const float4*>(dst + 4);
              } else {
                r0 = make_float4(0.f, 0.f, 0.f, 0.f);
                r1 = r0;
              }
              r0.x += f[0]; r0.y += f[1]; r0.z += f[2]; r0.w += f[3];
              r1.x += f[4]; r1.y += f[5]; r1.z += f[6]; r1.w += f[7];
              *reinterpret_cast<float4*>(dst) = r0;
              *reinterpret_cast<float4*>(dst + 4) = r1;
            }
          }
        }
      }
    }
}

// Pooling epilogue (EPI_BF16_HILO_POOL): TMEM lane = X row `row` (a gathered source token), G consecutive rows = the taps of one
// pooled token.  Each lane scales its act(acc + bias) by its tap weight, the G lanes of a group all-reduce by xor-shuffles
// (G divides 32 and groups never straddle a warp since tiles start at multiples of 32 rows), and lane j of the group writes
// the 32/G columns [j * 32/G, ...) of the pooled row as bf16 hi and lo: per-thread global stores, the output may be peer memory.
template <int ACT>
__device__ __forceinline__ void epilogue_pool(const GemmKernelParams& p, uint32_t t_row, int row, int y0, int c_begin, int c_end) {
  const int G = p.pool_group;
  const float wgt = row < p.x_rows ? __ldg(p.row_w + (p.row_w_period > 0 ? row % p.row_w_period : row)) : 0.f;
  const int j = (int)lane_id() & (G - 1);
  const int per = 32 / G;                       // columns this lane stores per 32-column chunk
  const long long orow = row / G;
  __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo;
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    if (y0 + c0 >= p.y_rows) break;  // warp-uniform
    uint32_t v[32];
    tmem_ld_32x32b_x32(t_row + c0, v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int n = y0 + c0 + i;
      const float b = (p.bias != nullptr && n < p.y_rows) ? __ldg(p.bias + n) : 0.f;
      f[i] = wgt * apply_act<ACT>(__uint_as_float(v[i]) + b);
    }
    for (int o = 1; o < G; o <<= 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] += __shfl_xor_sync(0xffffffffu, f[i], o);
    }
    if (row < p.x_rows) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i / per == j) {                     // resolved per lane; 32/G stores of 2 B hi + 2 B lo (G = 4: 8 columns)
          const int n = y0 + c0 + i;
          if (n < p.y_rows) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(f[i]);
            outp[n] = hi;
            outp[p.y_rows + n] = __float2bfloat16_rn(f[i] - __bfloat162float(hi));
          }
        }
      }
    }
  }
}

template <int BN, bool DUAL, int EPI, int ACT, bool HILO = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2,
                    const __grid_constant__ CUtensorMap tmY, const GemmKernelParams p) {
  using Cfg = GemmCfg<BN, DUAL, HILO>;
  constexpr int STAGES = Cfg::STAGES;
  // Tile order: the Y tile index varies fastest.  Swap-AB (weights on X): the CTAs that share a weight tile run at the
  // same time, so the second reader hits L2 instead of streaming the weights twice from HBM.  Normal orientation
  // (activations on X): all weight tiles of an activation row block are consumed together, so a large A (fc2: 200 MB,
  // more than L2) is streamed once instead of once per weight tile.
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int total_tiles = p.x_tiles * p.y_tiles * p.k_splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    if constexpr (DUAL) tma_prefetch_desc(&tmX2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 32 * GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  griddep_launch_dependents();
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      // k-block iterator over this CTA's tiles
      int tile = blockIdx.x, kb = 0, kb1 = 0, xt = 0, yt = 0;
      auto load_tile = [&]() {
        if (tile < total_tiles) {
          const int ks = tile % p.k_splits;
          const int rest = tile / p.k_splits;
          yt = rest % p.y_tiles; xt = rest / p.y_tiles;
          kb = ks * p.kb_per_split;
          kb1 = min(p.kb_total, kb + p.kb_per_split);
        }
      };
      auto advance = [&]() {
        if (++kb >= kb1) { tile += gridDim.x; load_tile(); }
      };
      load_tile();
      auto load_x = [&](const CUtensorMap* m, uint32_t dst, uint32_t bar) {
        if (p.x_blocked) tma_load_4d(dst, m, bar, 0, 0, kb, xt);
        else tma_load_2d(dst, m, bar, kb * BK, xt * BM);
      };
      uint32_t stage = 0, phase = 0;
      if (p.pdl_prefetch_x) {
        // Programmatic dependent launch: the X operand (weights) does not depend on the preceding kernels, so the first
        // ring of stages streams weight tiles from HBM while those kernels are still running; the Y operand
        // (activations) is loaded only after griddepcontrol.wait.
        int pre_kb[STAGES], pre_yt[STAGES];
        int n_pre = 0;
        while (tile < total_tiles && n_pre < STAGES) {
          const uint32_t sX = smem_base + n_pre * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(full_bar(n_pre), Cfg::STAGE_BYTES);
          load_x(&tmX, sX, full_bar(n_pre));
          if constexpr (DUAL) load_x(&tmX2, sX + Cfg::X_BYTES, full_bar(n_pre));
          pre_kb[n_pre] = kb; pre_yt[n_pre] = yt;
          ++n_pre;
          advance();
        }
        griddep_wait();
        for (int i = 0; i < n_pre; ++i) {
          const uint32_t sY = smem_base + i * Cfg::STAGE_BYTES + Cfg::X_BYTES * (DUAL ? 2 : 1);
          tma_load_2d(sY, &tmY, full_bar(i), pre_kb[i] * BK, pre_yt[i] * BN);
          if constexpr (HILO) tma_load_2d(sY + Cfg::Y_HALF, &tmY, full_bar(i), p.y_lo_off + pre_kb[i] * BK, pre_yt[i] * BN);
        }
        if (n_pre == STAGES) { stage = 0; phase = 1; } else { stage = n_pre; phase = 0; }
      } else {
        griddep_wait();
      }
      while (tile < total_tiles) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sX = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t sY = sX + Cfg::X_BYTES * (DUAL ? 2 : 1);
        mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
        load_x(&tmX, sX, full_bar(stage));
        if constexpr (DUAL) load_x(&tmX2, sX + Cfg::X_BYTES, full_bar(stage));
        tma_load_2d(sY, &tmY, full_bar(stage), kb * BK, yt * BN);
        if constexpr (HILO) tma_load_2d(sY + Cfg::Y_HALF, &tmY, full_bar(stage), p.y_lo_off + kb * BK, yt * BN);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        advance();
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ks = tile % p.k_splits;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sX = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sY = sX + Cfg::X_BYTES * (DUAL ? 2 : 1);
          const uint64_t dX = umma_desc_k_sw128(sX);
          const uint64_t dY = umma_desc_k_sw128(sY);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
            // advancing K inside the 128-B swizzle atom = +32 B on the start address (encoded >> 4)
            umma_f16(d_tmem, dX + 2u * k, dY + 2u * k, idesc, accum);
            if constexpr (HILO) umma_f16(d_tmem, dX + 2u * k, umma_desc_k_sw128(sY + Cfg::Y_HALF) + 2u * k, idesc, 1u);
            if constexpr (DUAL) {
              const uint64_t dX2 = umma_desc_k_sw128(sX + Cfg::X_BYTES);
              umma_f16(d_tmem + BN, dX2 + 2u * k, dY + 2u * k, idesc, accum);
              if constexpr (HILO) umma_f16(d_tmem + BN, dX2 + 2u * k, umma_desc_k_sw128(sY + Cfg::Y_HALF) + 2u * k, idesc, 1u);
            }
          }
          umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (TMEM -> registers -> global) =====================
    griddep_wait();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;                      // which half of the columns this warp drains
    const int c_begin = chalf * (BN / 2), c_end = c_begin + BN / 2;
    const int lane_row = q * 32 + (int)lane_id();
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ks = tile % p.k_splits;
      const int rest = tile / p.k_splits;
      const int xt = rest / p.y_tiles, yt = rest % p.y_tiles;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_COLS;
      const int xi = xt * BM + lane_row;  // index along X rows (TMEM lane)
      const int y0 = yt * BN;             // first index along Y rows (TMEM column 0)

      if constexpr (EPI == EPI_BF16_HILO_POOL) {
        epilogue_pool<ACT>(p, t_row, xi, y0, c_begin, c_end);
      } else if constexpr (EPI == EPI_BF16 || EPI == EPI_RESID_F32 || EPI == EPI_F32 || EPI == EPI_BF16_HILO) {
        epilogue_normal<EPI, ACT>(p, t_row, xi, y0, c_begin, c_end);
      } else if constexpr (EPI == EPI_T_F32) {
        // swap-AB: lane = n (weight row), column = token. Lanes of a warp write 32 consecutive n -> 128-B stores.
        const bool n_ok = xi < p.x_rows;
        float* plane = reinterpret_cast<float*>(p.out) + (long long)ks * p.split_stride;
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          if (y0 + c0 >= p.y_rows) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c0, v);
          tmem_ld_wait();
          if (n_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int tok = y0 + c0 + j;
              if (tok < p.y_rows) plane[(long long)tok * p.ldo + xi] = __uint_as_float(v[j]);
            }
          }
        }
      } else if constexpr (EPI == EPI_T_SWIGLU_IL) {
        // lane = interleaved weight row: even lanes hold gate_j, odd lanes up_j (j = xi / 2).  Each pair exchanges its
        // values by shuffle, computes silu(gate) * up, then pairs of pairs pack two outputs: lanes with (lane & 3) == 0
        // store 4 bytes, i.e. a warp writes 32 contiguous bytes (16 output features) per token.
        const bool n_ok = xi < p.x_rows;            // x_rows is even, so a pair is in or out together
        const bool odd = (lane_row & 1) != 0;
        const bool writer = (lane_row & 3) == 0;
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          if (y0 + c0 >= p.y_rows) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c0, v);
          tmem_ld_wait();
          if (p.raw_out != nullptr && y0 + c0 + 32 > p.raw_from && n_ok) {   // warp-uniform: only the chunks holding appended rows
#pragma unroll
            for (int j = 0; j < 32; ++j) {                   // lanes hold consecutive interleaved rows: a 128-B store per token
              const int tk = y0 + c0 + j;
              if (tk >= p.raw_from && tk < p.y_rows) p.raw_out[(long long)(tk - p.raw_from) * p.ld_raw + xi] = __uint_as_float(v[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float mine = __uint_as_float(v[j]);
            const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
            const float gv = odd ? other : mine, uv = odd ? mine : other;
            const float r = gv / (1.0f + __expf(-gv)) * uv;
            const float r2 = __shfl_xor_sync(0xffffffffu, r, 2);        // the next pair's result
            const int tok = y0 + c0 + j;
            if (writer && n_ok && tok < p.y_rows) {
              __nv_bfloat162 o = __floats2bfloat162_rn(r, r2);
              if (xi + 2 < p.x_rows) *reinterpret_cast<__nv_bfloat162*>(outp + (long long)tok * p.ldo + (xi >> 1)) = o;
              else outp[(long long)tok * p.ldo + (xi >> 1)] = o.x;   // last pair of an odd pair count
            }
          }
        }
      } else if constexpr (EPI == EPI_T_SWIGLU) {
        const bool n_ok = xi < p.x_rows;
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
          if (y0 + c0 >= p.y_rows) break;
          uint32_t g[16], u[16];
          tmem_ld_32x32b_x16(t_row + c0, g);
          tmem_ld_32x32b_x16(t_row + BN + c0, u);
          tmem_ld_wait();
          if (n_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int tok = y0 + c0 + j;
              if (tok < p.y_rows) {
                const float gv = __uint_as_float(g[j]);
                const float uv = __uint_as_float(u[j]);
                if (p.raw_out != nullptr && tok >= p.raw_from)
                  *reinterpret_cast<float2*>(p.raw_out + (long long)(tok - p.raw_from) * p.ld_raw + 2 * xi) = make_float2(gv, uv);
                const float s = gv / (1.0f + __expf(-gv));
                const __nv_bfloat16 hi = __float2bfloat16_rn(s * uv);
                outp[(long long)tok * p.ldo + xi] = hi;
                if (p.out_hilo) outp[(long long)tok * p.ldo + p.x_rows + xi] = __float2bfloat16_rn(s * uv - __bfloat162float(hi));
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// Epilogue through shared memory and TMA: the warp's 32 accumulator rows x 32 (fp32) / 64 (bf16) columns are staged in a
// 128-B-swizzled 4 KB tile and written by ONE bulk tensor store (or an fp32 reduce-add performed in L2 for the residual
// stream).  Replaces 16-B per-thread stores to 32 different rows per instruction, whose request stream kept the
// SM<->crossbar port 60% busy (ncu l1tex__m_l1tex2xbar_req_cycles_active) and starved the TMA operand loads.
// Rows / columns beyond the tensor are clipped by the TMA unit.
__device__ __forceinline__ uint32_t stage_sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

template <int EPI, int ACT>
__device__ __forceinline__ void epilogue_normal_tma(const GemmKernelParams& p, const CUtensorMap* tmOut, uint32_t t_row, int m_warp,
                                                    int y0, int c_begin, int c_end, uint32_t stage, int plane) {
  const int lane = (int)lane_id();
  if constexpr (EPI == EPI_SWIGLU_PAIR) {
    // 128 accumulator columns = 64 (gate, up) pairs -> 64 bf16 outputs = one 128-B staging row
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 128) {
      if (y0 + c0 >= p.y_rows) break;
      uint32_t o[32];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c0 + 32 * h, v);
        tmem_ld_wait();
        if (p.raw_out != nullptr) {     // appended precise rows: raw (gate, up) pairs, this lane's row, 32 columns
          const int row = m_warp + lane;
          if (row >= p.raw_from && row < p.x_rows) {
            float4* dst = reinterpret_cast<float4*>(p.raw_out + (long long)(row - p.raw_from) * p.ld_raw + y0 + c0 + 32 * h);
#pragma unroll
            for (int g = 0; g < 8; ++g)
              dst[g] = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
          }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {   // 4 accumulator columns -> 2 outputs -> one packed word
          const float g0 = __uint_as_float(v[4 * g]), u0 = __uint_as_float(v[4 * g + 1]);
          const float g1 = __uint_as_float(v[4 * g + 2]), u1 = __uint_as_float(v[4 * g + 3]);
          o[8 * h + g] = pack_bf16(g0 / (1.0f + __expf(-g0)) * u0, g1 / (1.0f + __expf(-g1)) * u1);
        }
      }
      if (lane == 0) bulk_wait_group_read0();
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + stage_sw128_off(lane, c)), "r"(o[4 * c]), "r"(o[4 * c + 1]),
                     "r"(o[4 * c + 2]), "r"(o[4 * c + 3]) : "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmOut, stage, (y0 + c0) >> 1, m_warp);
        bulk_commit_group();
      }
    }
    return;
  }
  if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_F32) {
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      if (y0 + c0 >= p.y_rows) break;
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c0, v);
      tmem_ld_wait();
      if (lane == 0) bulk_wait_group_read0();     // the previous store has finished reading the staging tile
      __syncwarp();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int n = y0 + c0 + g * 4;
        float4 f = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
        if (p.bias != nullptr && n < p.y_rows) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          f.x += b.x; f.y += b.y; f.z += b.z; f.w += b.w;
        }
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage + stage_sw128_off(lane, g)), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if constexpr (EPI == EPI_RESID_F32) tma_reduce_add_2d(tmOut, stage, y0 + c0, m_warp);
        else tma_store_3d(tmOut, stage, y0 + c0, m_warp, plane);
        bulk_commit_group();
      }
    }
  } else {
    // bf16 outputs: 64 columns (128 B per row) per store; EPI_BF16_HILO writes a second tile p.y_rows columns further right
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 64) {
      if (y0 + c0 >= p.y_rows) break;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c0 + 32 * h, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int n = y0 + c0 + 32 * h + g * 4;
          float f[4] = {__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3])};
          if (p.bias != nullptr && n < p.y_rows) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            f[0] += b.x; f[1] += b.y; f[2] += b.z; f[3] += b.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) f[j] = apply_act<ACT>(f[j]);
          hi[16 * h + 2 * g] = pack_bf16(f[0], f[1]);
          hi[16 * h + 2 * g + 1] = pack_bf16(f[2], f[3]);
          if constexpr (EPI == EPI_BF16_HILO) {
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[16 * h + 2 * g]));
            const float2 b2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[16 * h + 2 * g + 1]));
            lo[16 * h + 2 * g] = pack_bf16(f[0] - a.x, f[1] - a.y);
            lo[16 * h + 2 * g + 1] = pack_bf16(f[2] - b2.x, f[3] - b2.y);
          }
        }
      }
      constexpr int NPASS = (EPI == EPI_BF16_HILO) ? 2 : 1;
#pragma unroll
      for (int pass = 0; pass < NPASS; ++pass) {
        if (lane == 0) bulk_wait_group_read0();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t* src = pass == 0 ? &hi[4 * c] : &lo[4 * c];
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + stage_sw128_off(lane, c)), "r"(src[0]), "r"(src[1]), "r"(src[2]),
                       "r"(src[3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmOut, stage, y0 + c0 + pass * p.y_rows, m_warp);
          bulk_commit_group();
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for the large-M GEMMs of the ViT / projector: two CTAs of a cluster (one TPC) share
// one UMMA of 256 x BN2 x 16.  Each CTA stages its own 128 rows of A and HALF of the B tile (BN2/2 rows), so the
// shared-memory traffic per MMA drops by a third compared with the single-CTA kernel, whose tensor pipe ncu shows only
// 63% active at 128 x 256 tiles.  The leader CTA (rank 0) issues the MMAs; both CTAs run a TMA producer (completion bytes
// are credited to the LEADER's full barrier) and their own epilogue over their 128 accumulator lanes; tcgen05.commit
// multicasts the "slot free" / "accumulator ready" arrivals to both CTAs.
// ------------------------------------------------------------------------------------------------------------
template <int BN2>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BM * BK * 2;                 // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN2 / 2) * BK * 2;          // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 512;                       // 2 accumulators of BN2 columns
  static constexpr int BAR_BYTES = 1024;
  static constexpr int EPI_STAGE_BYTES = 4096 * GEMM_EPI_WARPS;   // one swizzled 32-row x 128-B staging tile per epilogue warp
  static constexpr int STAGES_RAW = (SMEM_BUDGET - 1024 - BAR_BYTES - EPI_STAGE_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + BAR_BYTES + 1024;
  static_assert(2 * BN2 <= 512 && BN2 % 32 == 0, "accumulator columns");
};

template <int BN2, int EPI, int ACT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmOut, const GemmKernelParams p) {
  using Cfg = Gemm2Cfg<BN2>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_stage_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = epi_stage_base + Cfg::EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int m_tiles = (p.x_rows + 2 * BM - 1) / (2 * BM);
  const int n_tiles = (p.y_rows + BN2 - 1) / BN2;
  const int total_tiles = m_tiles * n_tiles * p.k_splits;   // split-K (EPI_F32 only): plane ks of a 3-D output map
  (void)m_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * 32 * GEMM_EPI_WARPS);   // the epilogue threads of BOTH CTAs arrive on the leader's barrier
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        const int ks = tile % p.k_splits, rest = tile / p.k_splits;
        const int nt = rest % n_tiles, mt = rest / n_tiles;   // weight tiles fastest: an A row block is read once from HBM
        const int kb_end = min(p.kb_total, (ks + 1) * p.kb_per_split);
        for (int kb = ks * p.kb_per_split; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          tma_load_2d_2cta(sA, &tmA, lead_full, kb * BK, mt * 2 * BM + (int)rank * BM);
          tma_load_2d_2cta(sB, &tmB, lead_full, kb * BK, nt * BN2 + (int)rank * (BN2 / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, single thread) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN2);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN2;
        const int ks = tile % p.k_splits;
        const int kb0 = ks * p.kb_per_split, kb_end = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t dA = umma_desc_k_sw128(sA);
          const uint64_t dB = umma_desc_k_sw128(sA + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_f16_2cta(d_tmem, dA + 2u * k, dB + 2u * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_2cta(empty_bar(stage), 0x3);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2cta(tfull_bar(acc), 0x3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (both CTAs, own 128 accumulator lanes) =====================
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int c_begin = chalf * (BN2 / 2), c_end = c_begin + BN2 / 2;
    const int lane_row = q * 32 + (int)lane_id();
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      const int ks = tile % p.k_splits, rest = tile / p.k_splits;
      const int nt = rest % n_tiles, mt = rest / n_tiles;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN2;
      const int m_warp = mt * 2 * BM + (int)rank * BM + q * 32;   // first output row of this warp's 32 lanes
      if constexpr (EPI == EPI_BF16_HILO_POOL) epilogue_pool<ACT>(p, t_row, m_warp + (int)lane_id(), nt * BN2, c_begin, c_end);
      else epilogue_normal_tma<EPI, ACT>(p, &tmOut, t_row, m_warp, nt * BN2, c_begin, c_end, epi_stage_base + (warp - 2) * 4096, ks);
      tc_fence_before();
      mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane_id() == 0) bulk_wait_group0();   // the staging tiles must outlive the last bulk stores
    (void)lane_row;
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA may free its TMEM / leave while the pair's MMAs or remote arrivals are in flight
  if (warp == 1) tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// Host side: tensor-map cache and dispatch
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapKey {
  const void* ptr; int rows; int K; long long ld; int box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && K == o.K && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= (size_t)k.K * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (size_t)k.ld * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    h ^= (size_t)k.box_rows + (h << 6) + (h >> 2);
    return h;
  }
};

struct GemmContext {
  int device = 0;
  int num_sms = 148;
  PFN_encodeTiled encode = nullptr;
  std::mutex mu;
  std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
};

GemmContext* gemm_context_create(int device) {
  auto* c = new GemmContext();
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    g_gemm_err = "cudaGetDeviceProperties failed";
    delete c;
    return nullptr;
  }
  c->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
    g_gemm_err = "cuTensorMapEncodeTiled entry point not found";
    delete c;
    return nullptr;
  }
  c->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return c;
}
void gemm_context_destroy(GemmContext* c) { delete c; }

static int get_map(GemmContext* c, const void* ptr, int rows, int K, long long ld, int box_rows, CUtensorMap* out) {
  MapKey key{ptr, rows, K, ld, box_rows};
  {
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->maps.find(key);
    if (it != c->maps.end()) { *out = it->second; return 0; }
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0) {
    g_gemm_err = "gemm operand must be 16-B aligned with a 16-B aligned row stride";
    return -2;
  }
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = c->encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%lld box=%d", (int)r, rows, K, ld, box_rows);
    g_gemm_err = buf;
    return -3;
  }
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->maps.size() > 65536) c->maps.clear();
    c->maps.emplace(key, m);
  }
  *out = m;
  return 0;
}

// Output tensor map for the TMA-store epilogue: [rows, width] fp32 (box 32 x 32) or bf16 (box 64 x 32), 128-B swizzle.
static int get_out_map(GemmContext* c, const void* ptr, int rows, int width, long long ld, bool f32, CUtensorMap* out, int planes = 0,
                       long long plane_stride = 0) {
  MapKey key{ptr, rows, width, ld, (f32 ? -2 : -3) - 16 * planes};
  {
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->maps.find(key);
    if (it != c->maps.end()) { *out = it->second; return 0; }
  }
  const int esz = f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * esz) % 16 != 0) {
    g_gemm_err = "gemm output must be 16-B aligned with a 16-B aligned row stride";
    return -2;
  }
  CUtensorMap m;
  const int rank = planes > 0 ? 3 : 2;   // planes > 0: split-K partial planes [planes][rows][width]
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)(planes > 0 ? planes : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * esz, (cuuint64_t)plane_stride * esz};
  cuuint32_t box[3] = {(cuuint32_t)(f32 ? 32 : 64), 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = c->encode(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(out) failed (%d) rows=%d width=%d ld=%lld", (int)r, rows, width, ld);
    g_gemm_err = buf;
    return -3;
  }
  {
    std::lock_guard<std::mutex> lk(c->mu);
    c->maps.emplace(key, m);
  }
  *out = m;
  return 0;
}

// 4-D map over tile-blocked weights [rows/128][K/64][128][64]: box {64,128,1,1} = one contiguous 16 KB tile.
static int get_map_blocked(GemmContext* c, const void* ptr, int rows, int K, CUtensorMap* out) {
  MapKey key{ptr, rows, K, -1, -1};
  {
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->maps.find(key);
    if (it != c->maps.end()) { *out = it->second; return 0; }
  }
  if (rows % BM != 0 || K % BK != 0 || (reinterpret_cast<uintptr_t>(ptr) & 127) != 0) {
    g_gemm_err = "blocked weights need rows % 128 == 0, K % 64 == 0 and 128-B alignment";
    return -2;
  }
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)BK, (cuuint64_t)BM, (cuuint64_t)(K / BK), (cuuint64_t)(rows / BM)};
  cuuint64_t strides[3] = {(cuuint64_t)BK * 2, (cuuint64_t)BK * BM * 2, (cuuint64_t)BK * BM * 2 * (cuuint64_t)(K / BK)};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)BM, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = c->encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(blocked) failed (%d) rows=%d K=%d", (int)r, rows, K);
    g_gemm_err = buf;
    return -3;
  }
  {
    std::lock_guard<std::mutex> lk(c->mu);
    c->maps.emplace(key, m);
  }
  *out = m;
  return 0;
}

int gemm_effective_splits(int K, int k_splits) {
  const int kb_total = (K + BK - 1) / BK;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > kb_total) k_splits = kb_total;
  const int per = (kb_total + k_splits - 1) / k_splits;
  return (kb_total + per - 1) / per;
}

template <int BN, bool DUAL, int EPI, int ACT, bool HILO = false>
static int launch_cfg(GemmContext* c, const GemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, DUAL, HILO>;
  auto kern = gemm_tcgen05_kernel<BN, DUAL, EPI, ACT, HILO>;
  if (HILO && (a.K % BK != 0 || a.ldy < 2 * (int64_t)a.K)) { g_gemm_err = "hi/lo Y operand needs K % 64 == 0 and ldy >= 2K"; return -2; }
  static bool attr_set[kMaxDevices] = {};  // per instantiation and per device (the attribute is per device)
  if (!attr_set[c->device & (kMaxDevices - 1)]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { g_gemm_err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return -4; }
    attr_set[c->device & (kMaxDevices - 1)] = true;
  }
  CUtensorMap tmX, tmX2, tmY;
  int rc;
  if (a.x_blocked) {
    if ((rc = get_map_blocked(c, a.X, a.x_rows, a.K, &tmX)) != 0) return rc;
    if (DUAL) { if ((rc = get_map_blocked(c, a.X2, a.x_rows, a.K, &tmX2)) != 0) return rc; }
    else tmX2 = tmX;
  } else {
    if ((rc = get_map(c, a.X, a.x_rows, a.K, a.ldx, BM, &tmX)) != 0) return rc;
    if (DUAL) { if ((rc = get_map(c, a.X2, a.x_rows, a.K, a.ldx, BM, &tmX2)) != 0) return rc; }
    else tmX2 = tmX;
  }
  if ((rc = get_map(c, a.Y, a.y_rows, HILO ? 2 * a.K : a.K, a.ldy, BN, &tmY)) != 0) return rc;

  GemmKernelParams p;
  p.y_lo_off = a.K; p.out_hilo = a.out_hilo; p.row_w = a.row_w; p.pool_group = a.pool_group; p.row_w_period = a.row_w_period;
  p.raw_out = a.raw_out; p.raw_from = a.raw_from; p.ld_raw = a.ld_raw;
  p.x_rows = a.x_rows; p.y_rows = a.y_rows;
  p.x_tiles = (a.x_rows + BM - 1) / BM;
  p.y_tiles = (a.y_rows + BN - 1) / BN;
  p.kb_total = (a.K + BK - 1) / BK;
  const int splits = (EPI == EPI_T_F32) ? gemm_effective_splits(a.K, a.k_splits) : 1;
  p.k_splits = splits;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.bias = a.bias; p.out = a.out; p.ldo = a.ldo; p.split_stride = a.split_stride;
  const long long tiles = (long long)p.x_tiles * p.y_tiles * p.k_splits;
  int max_ctas = a.max_ctas > 0 ? a.max_ctas : c->num_sms;
  const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  if (grid <= 0) return 0;
  p.x_blocked = a.x_blocked;
  p.pdl_prefetch_x = (g_use_pdl && (EPI == EPI_T_F32 || EPI == EPI_T_SWIGLU || EPI == EPI_T_SWIGLU_IL)) ? 1 : 0;
  cudaError_t e = launch_k(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmX, tmX2, tmY, p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { g_gemm_err = std::string("gemm launch: ") + cudaGetErrorString(e); return -5; }
  return 0;
}

int g_gemm_use_2cta = 1;   // CTA-pair kernel for the large-M normal-orientation GEMMs (mmd_set_gemm_2cta)

template <int BN2, int EPI, int ACT>
static int launch_2cta(GemmContext* c, const GemmArgs& a, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN2>;
  auto kern = gemm_tcgen05_2cta_kernel<BN2, EPI, ACT>;
  static bool attr_set[kMaxDevices] = {};
  if (!attr_set[c->device & (kMaxDevices - 1)]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { g_gemm_err = std::string("cudaFuncSetAttribute(2cta): ") + cudaGetErrorString(e); return -4; }
    attr_set[c->device & (kMaxDevices - 1)] = true;
  }
  CUtensorMap tmA, tmB, tmOut;
  int rc;
  if ((rc = get_map(c, a.X, a.x_rows, a.K, a.ldx, BM, &tmA)) != 0) return rc;
  if ((rc = get_map(c, a.Y, a.y_rows, a.K, a.ldy, BN2 / 2, &tmB)) != 0) return rc;
  const int splits = (EPI == EPI_F32) ? gemm_effective_splits(a.K, a.k_splits) : 1;
  {
    const bool f32 = (EPI == EPI_RESID_F32 || EPI == EPI_F32);
    const int width = (EPI == EPI_BF16_HILO || EPI == EPI_BF16_HILO_POOL) ? a.y_rows * 2 : (EPI == EPI_SWIGLU_PAIR ? a.y_rows / 2 : a.y_rows);
    const int out_rows = EPI == EPI_BF16_HILO_POOL ? a.x_rows / a.pool_group : a.x_rows;   // (the pooling epilogue stores directly; map unused)
    // EPI_F32 always uses the 3-D form (plane coordinate = split index; one plane when there is no split-K)
    if ((rc = get_out_map(c, a.out, out_rows, width, a.ldo, f32, &tmOut, EPI == EPI_F32 ? splits : 0,
                          splits > 1 ? a.split_stride : (long long)a.x_rows * a.ldo)) != 0) return rc;
  }
  GemmKernelParams p;
  p.x_rows = a.x_rows; p.y_rows = a.y_rows;
  p.x_tiles = (a.x_rows + 2 * BM - 1) / (2 * BM);
  p.y_tiles = (a.y_rows + BN2 - 1) / BN2;
  p.kb_total = (a.K + BK - 1) / BK;
  p.k_splits = splits; p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.bias = a.bias; p.out = a.out; p.ldo = a.ldo; p.split_stride = 0; p.pdl_prefetch_x = 0; p.x_blocked = 0;
  p.y_lo_off = 0; p.out_hilo = 0; p.row_w = a.row_w; p.pool_group = a.pool_group; p.row_w_period = a.row_w_period;
  p.raw_out = a.raw_out; p.raw_from = a.raw_from; p.ld_raw = a.ld_raw;
  const long long tiles = (long long)p.x_tiles * p.y_tiles * splits;
  const int max_clusters = (a.max_ctas > 0 ? a.max_ctas : c->num_sms) / 2;
  const int clusters = (int)(tiles < max_clusters ? tiles : max_clusters);
  if (clusters <= 0) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { g_gemm_err = std::string("gemm 2cta launch: ") + cudaGetErrorString(e); return -5; }
  return 0;
}

template <int EPI, int ACT>
static int launch_normal(GemmContext* c, const GemmArgs& a, cudaStream_t s) {
  if (a.y_rows % 8 != 0) { g_gemm_err = "normal-orientation GEMM needs N % 8 == 0"; return -2; }
  if (g_gemm_use_2cta && !a.force_1cta && (a.x_rows >= 1024 || a.force_2cta) && a.y_rows >= 192) {
    // N = 1152 (out_proj / fc2 / patch embed) tiles exactly by 192
    if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_F32) {
      static const bool no192 = getenv("MMD_NO_BN192") != nullptr;
      if (!no192 && a.y_rows % 192 == 0 && a.y_rows % 256 != 0 && a.y_rows <= 1536) return launch_2cta<192, EPI, ACT>(c, a, s);
    }
    return launch_2cta<256, EPI, ACT>(c, a, s);
  }
  if (a.y_rows <= 128) return launch_cfg<128, false, EPI, ACT>(c, a, s);
  // N = 1152 (SigLIP out_proj / fc2 / patch embed) tiles exactly by 192 and wastes 10% of the MMAs with 256-wide tiles
  if (a.y_rows % 192 == 0 && a.y_rows % 256 != 0 && a.y_rows <= 1536) return launch_cfg<192, false, EPI, ACT>(c, a, s);
  return launch_cfg<256, false, EPI, ACT>(c, a, s);
}

int gemm_launch(GemmContext* c, const GemmArgs& a, cudaStream_t s) {
  if (c == nullptr) { g_gemm_err = "null gemm context"; return -1; }
  if (a.X == nullptr || a.Y == nullptr || a.out == nullptr || a.x_rows <= 0 || a.y_rows <= 0 || a.K <= 0) {
    g_gemm_err = "gemm: null pointer or non-positive shape";
    return -2;
  }
  if (a.K % 8 != 0) { g_gemm_err = "gemm: K must be a multiple of 8 (16-B rows for TMA)"; return -2; }
  switch (a.epi) {
    case EPI_BF16:
      if (a.act == ACT_NONE) return launch_normal<EPI_BF16, ACT_NONE>(c, a, s);
      if (a.act == ACT_GELU_TANH) return launch_normal<EPI_BF16, ACT_GELU_TANH>(c, a, s);
      if (a.act == ACT_GELU_ERF) return launch_normal<EPI_BF16, ACT_GELU_ERF>(c, a, s);
      break;
    case EPI_BF16_HILO:
      if (a.act == ACT_GELU_ERF) return launch_normal<EPI_BF16_HILO, ACT_GELU_ERF>(c, a, s);
      if (a.act == ACT_NONE) return launch_normal<EPI_BF16_HILO, ACT_NONE>(c, a, s);
      break;
    case EPI_BF16_HILO_POOL:
      if (a.row_w == nullptr || (a.pool_group != 4 && a.pool_group != 16) || a.x_rows % a.pool_group != 0 || a.ldo < 2 * (int64_t)a.y_rows) {
        g_gemm_err = "pooling epilogue needs row_w, pool_group 4 or 16 dividing x_rows, ldo >= 2N";
        return -2;
      }
      if (a.act == ACT_GELU_ERF) return launch_normal<EPI_BF16_HILO_POOL, ACT_GELU_ERF>(c, a, s);
      if (a.act == ACT_NONE) return launch_normal<EPI_BF16_HILO_POOL, ACT_NONE>(c, a, s);
      break;
    case EPI_SWIGLU_PAIR:
      if (a.y_rows % 256 != 0) { g_gemm_err = "swiglu-pair gemm needs N % 256 == 0"; return -2; }
      return launch_2cta<256, EPI_SWIGLU_PAIR, ACT_NONE>(c, a, s);
    case EPI_RESID_F32: return launch_normal<EPI_RESID_F32, ACT_NONE>(c, a, s);
    case EPI_F32: return launch_normal<EPI_F32, ACT_NONE>(c, a, s);
    case EPI_T_F32:
      if (a.y_hilo) {
        if (a.y_rows <= 64) return launch_cfg<64, false, EPI_T_F32, ACT_NONE, true>(c, a, s);
        if (a.y_rows <= 128) return launch_cfg<128, false, EPI_T_F32, ACT_NONE, true>(c, a, s);
        g_gemm_err = "hi/lo activations: at most 128 rows per launch";
        return -2;
      }
      if (a.y_rows <= 64) return launch_cfg<64, false, EPI_T_F32, ACT_NONE>(c, a, s);
      if (a.y_rows <= 128) return launch_cfg<128, false, EPI_T_F32, ACT_NONE>(c, a, s);
      return launch_cfg<256, false, EPI_T_F32, ACT_NONE>(c, a, s);
    case EPI_T_SWIGLU_IL:
      if (a.x_rows % 4 != 0 || a.ldo % 2 != 0) { g_gemm_err = "interleaved swiglu gemm needs x_rows % 4 == 0 and an even ldo"; return -2; }
      return launch_cfg<256, false, EPI_T_SWIGLU_IL, ACT_NONE>(c, a, s);
    case EPI_T_SWIGLU:
      if (a.X2 == nullptr) { g_gemm_err = "swiglu gemm needs X2"; return -2; }
      if (a.y_hilo) {
        if (a.y_rows <= 64) return launch_cfg<64, true, EPI_T_SWIGLU, ACT_NONE, true>(c, a, s);
        if (a.y_rows <= 128) return launch_cfg<128, true, EPI_T_SWIGLU, ACT_NONE, true>(c, a, s);
        g_gemm_err = "hi/lo activations: at most 128 rows per launch";
        return -2;
      }
      if (a.y_rows <= 64) return launch_cfg<64, true, EPI_T_SWIGLU, ACT_NONE>(c, a, s);
      return launch_cfg<128, true, EPI_T_SWIGLU, ACT_NONE>(c, a, s);
    default: break;
  }
  g_gemm_err = "gemm: unknown epilogue/activation";
  return -2;
}

}  // namespace mmd
