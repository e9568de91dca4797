// Host-side interface of the tcgen05 GEMM (see gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace mmd {

// Epilogue selector.
//  "Normal" orientation: X = activations [M,K] (TMEM lanes = rows m), Y = weights [N,K] (TMEM columns = n).
//  "T" (swap-AB) orientation: X = weights [N,K] (TMEM lanes = n), Y = activations [M,K] (columns = token m);
//  used when M is small so the 128-wide UMMA M dimension is filled by weight rows and HBM streaming of the
//  weights is spread over all SMs with split-K.
enum GemmEpi : int {
  EPI_BF16 = 0,      // out_bf16[m, n] = act(acc + bias[n])
  EPI_RESID_F32 = 1, // out_f32[m, n] += acc + bias[n]            (fp32 residual stream, in place)
  EPI_T_F32 = 2,     // out_f32[split][m, n] = acc                (split-K partial planes, no bias)
  EPI_T_SWIGLU = 3,  // out_bf16[m, n] = silu(acc_gate) * acc_up  (two X operands: gate rows, up rows)
  EPI_F32 = 4,       // out_f32[m, n] = acc + bias[n]
  EPI_SWIGLU_PAIR = 6, // normal orientation, weights interleaved (row 2j = gate_j, row 2j+1 = up_j): out_bf16[m, j] = silu(acc[2j]) * acc[2j+1]
  EPI_T_SWIGLU_IL = 7, // swap-AB, ONE operand with interleaved rows (2j = gate_j, 2j+1 = up_j): out_bf16[m, j] = silu(acc[2j]) * acc[2j+1];
                       // single accumulator, 256-token tiles (UMMA N = 256 instead of 2 x 128); out is [M, x_rows/2]
  EPI_BF16_HILO = 5, // v = act(acc + bias[n]); out_bf16[m, n] = hi = bf16(v); out_bf16[m, N + n] = bf16(v - hi)  (ldo >= 2N)
  // Spatial pooling in the epilogue: every group of `pool_group` (4 or 16) consecutive X rows is one pooled output row,
  // v[g, n] = sum_i row_w[g*G + i] * act(acc[g*G + i, n] + bias[n]) (fp32, combined across TMEM lanes by shuffles), written as
  // a hi+lo pair like EPI_BF16_HILO; out is [x_rows / G, 2N].  mm_projector.0 + GELU + the bilinear / average taps.
  EPI_BF16_HILO_POOL = 8,
};
enum GemmAct : int { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2 };

struct GemmArgs {
  const __nv_bfloat16* X = nullptr;   // [x_rows, K], row stride ldx (elements)
  const __nv_bfloat16* X2 = nullptr;  // second lane operand (EPI_T_SWIGLU: up_proj rows), same shape as X
  const __nv_bfloat16* Y = nullptr;   // [y_rows, K], row stride ldy
  int x_rows = 0, y_rows = 0, K = 0;
  // X (and X2) stored tile-blocked: [x_rows/128][K/64][128][64] bf16, i.e. every 128x64 operand tile is one contiguous
  // 16 KB burst in HBM (needs x_rows % 128 == 0 and K % 64 == 0).  Static weights are packed this way once at load.
  int x_blocked = 0;
  int64_t ldx = 0, ldy = 0;
  int epi = EPI_BF16, act = ACT_NONE;
  const float* bias = nullptr;        // may be null
  void* out = nullptr;
  int64_t ldo = 0;                    // output row stride (elements)
  int k_splits = 1;                   // only EPI_T_F32
  int64_t split_stride = 0;           // elements between partial planes
  int max_ctas = 0;                   // 0 = number of SMs
  int force_2cta = 0;                 // use the CTA-pair kernel even for M < 1024 (decoder passes with >= ~200 tokens)
  // swap-AB kernels (EPI_T_F32 / EPI_T_SWIGLU), at most 128 Y rows: Y is a bf16 hi+lo pair [hi | lo] of width 2K (ldy >= 2K);
  // each weight tile is multiplied with both halves into the same accumulator ("precise rows")
  const float* row_w = nullptr;       // EPI_BF16_HILO_POOL: one pooling weight per X row
  int pool_group = 0;                 // EPI_BF16_HILO_POOL: rows per pooled output (4 or 16; x_rows % pool_group == 0)
  int row_w_period = 0;               // > 0: row_w holds one period (a frame) and is indexed by row % period
  int force_1cta = 0;                 // single-CTA kernel with per-thread global stores (outputs in peer memory)
  // SwiGLU epilogues (EPI_T_SWIGLU, EPI_T_SWIGLU_IL, EPI_SWIGLU_PAIR): tokens >= raw_from are the appended hi / lo rows of the
  // "precise rows"; for them the raw fp32 pre-activations are written to raw_out[(token - raw_from) * ld_raw + 2j + {0: gate,
  // 1: up}] (the nonlinearity needs gate_hi + gate_lo first; swiglu_from_raw finishes them).  raw_out == nullptr: off.
  float* raw_out = nullptr;
  int raw_from = 0;
  int64_t ld_raw = 0;
  int y_hilo = 0;
  int out_hilo = 0;                   // EPI_T_SWIGLU: out row = [hi | lo], lo at column x_rows (ldo >= 2 * x_rows)
};

struct GemmContext;  // tensor-map cache + device properties
GemmContext* gemm_context_create(int device);
void gemm_context_destroy(GemmContext*);
// Returns 0 on success, negative on bad arguments / CUDA error; message via gemm_last_error().
int gemm_launch(GemmContext*, const GemmArgs&, cudaStream_t stream);
const char* gemm_last_error();
// Effective number of split-K planes the launch will write for (K, requested splits).
int gemm_effective_splits(int K, int k_splits);
extern int g_gemm_use_2cta;  // 1 (default): cta_group::2 kernel for the large-M normal-orientation GEMMs

}  // namespace mmd
