// Flash attention on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM) for sm_100a.
//
// One CTA owns TWO 128-row query tiles of the same head (ping-pong): while the softmax warps of one tile work on their
// scores, the tensor pipe computes the other tile's products.  Keys/values stream in 64-key tiles shared by both.
//
//   S = Q K^T   : UMMA 128 x 64 x 16, A = Q tile (smem, K-major), B = K tile (smem, K-major)  -> TMEM, 2 buffers per q tile
//   O += P V    : UMMA 128 x DH x 16, A = P tile read FROM TMEM (bf16 pairs written over the S buffer it came from),
//                 B = V tile (smem, MN-major: the same [key][d] image the loader writes for K, no transpose)
//                                                                                               -> TMEM, accumulated there
//   softmax     : 2 x 4 warps, ONE THREAD PER QUERY ROW (TMEM lane): tcgen05.ld the score row, running max/sum in
//                 registers with no shuffles, P stored back to TMEM with tcgen05.st (no shared-memory round trip, and
//                 no wait for the previous PV: the only loop-carried dependency left is S_j itself).
//                 O stays in TMEM across tiles; it is rescaled (tcgen05.ld / st) only when the running maximum grows by
//                 more than 2^8 ("lazy rescaling"), otherwise the stale maximum keeps serving as the exponent base.
//   loaders     : 4 warps, 16-B cp.async into the swizzled operand layout (zero fill for rows past the end and for the
//                 head-dim padding 72 -> 80), two tiles in flight, handed over with fence.proxy.async + mbarrier.
//   MMA         : one thread issues every tcgen05.mma and the tcgen05.commit that signals the mbarriers.
//
// Two front ends share the pipeline:
//   * vit:   non-causal attention over the 729 patch tokens of a frame, packed qkv [T*S, 3*H*72]
//            (TF:models/siglip/modeling_siglip.py:252-330), optional hi+lo output for the out-projection;
//   * paged: KV-append attention of the Qwen2 decoder over the paged KV pool (64-token pages = one key tile) with GQA row
//            stacking, bottom-right causal mask and split-KV partial outputs (TF:models/qwen2/modeling_qwen2.py:187-246).
#include "kernels.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <math.h>

namespace mmd {

namespace {

constexpr int TA_BM = 128;                 // query rows per tile = TMEM lanes
constexpr int TA_QT = 2;                   // query tiles per CTA
constexpr int TA_BN = 64;                  // keys per tile
constexpr int TA_QREGION = TA_BM * 128;    // one 64-column swizzled region of a 128-row operand (16 KB)
constexpr int TA_KREGION = TA_BN * 128;    // same for a 64-row operand (8 KB)
constexpr int TA_SOFTMAX_WARPS = 4 * TA_QT, TA_LOADER_WARPS = 4;
constexpr int TA_THREADS = 32 * (TA_SOFTMAX_WARPS + TA_LOADER_WARPS + 4);   // MMA warp + 3 idle warps: setmaxnreg works on whole warpgroups
constexpr int TA_TMEM_PER_Q = 256;         // S0 [0,64) S1 [64,128) O [128, 128+DHP)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
#ifdef MMD_SYNC_LOADERS   // diagnostic builds only: wait for the copies, fence on the WRITER side, then a plain arrive
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  fence_proxy_async_smem();
  mbar_arrive(bar);
#else
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// byte offset of 16-B chunk `c` (0..7) of row `r` inside a 128-B-swizzled K-major region (rows 128 B apart, 8-row
// groups 1024 B apart): exactly the image TMA writes with CU_TENSOR_MAP_SWIZZLE_128B.
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// A format: 1 = bf16, 0 = f16 (the P operand when it is produced by packed half-precision exponentials); B = V is bf16
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn_b(uint32_t M, uint32_t N, uint32_t a_bf16 = 1u) {
  return (1u << 4) | (a_bf16 << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Packed exponentials (MMD_EXP_F16X2, experiment, default OFF): ex2.approx.f16x2 / ex2.approx.ftz.bf16x2 would halve the
// MUFU work if the unit were packed, and the f16x2 result could be the PV product's A operand as it is (P as f16 read from
// TMEM, V stays bf16).  On sm_100a it is not: ptxas lowers either form to TWO scalar ops (MUFU.EX2.F16 Rd, Rs and
// MUFU.EX2.F16 Rd, Rs.H1) plus a PRMT, i.e. the same MUFU count as two fp32 ex2 and one more ALU op
// (cuobjdump of this file with -DMMD_EXP_F16X2=1: 64 MUFU.EX2.F16 per kernel for 32 packed ex2).  Kept for the record.
#ifndef MMD_EXP_F16X2
#define MMD_EXP_F16X2 0
#endif
#if MMD_EXP_F16X2
__device__ __forceinline__ uint32_t ex2_f16x2(float lo, float hi) {
  uint32_t h, r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(h));
  return r;
}
__device__ __forceinline__ uint32_t hadd2_u32(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ float h2_sum_f32(uint32_t h) {
  float lo, hi;
  asm("{ .reg .b16 l, u; mov.b32 {l, u}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, u; }" : "=f"(lo), "=f"(hi) : "r"(h));
  return lo + hi;
}
#endif
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 2^x on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, max relative error 7.5e-5, far below the bf16
// rounding of P): MUFU.EX2 runs at 16 lanes/clk/SM, so a share of every tile's exponentials is computed here instead
// and overlaps the MUFU stream.  Valid for x <= 127; anything below -125 is clamped (2^-125 rounds to nothing in O).
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;                 // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float f = x - (t - 12582912.f);           // [-0.5, 0.5]
  float p = fmaf(f, 0.0551714599f, 0.2426108569f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999281168f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// which of the 16 element pairs of a 32-key block go to the polynomial (3 of every 8)
#ifndef MMD_EXP_MODE
#define MMD_EXP_MODE 0
#endif
__device__ __forceinline__ constexpr bool exp_on_fma_pipe(int i) {
  return MMD_EXP_MODE == 2 ? true : (MMD_EXP_MODE == 0 ? false : ((i & 7) == 1 || (i & 7) == 4 || (i & 7) == 6));
}

// Diagnostic builds (-DMMD_ATTN_JITTER): pseudo-random sleeps in every role, to shake the hand-over protocol's timing.
#ifdef MMD_ATTN_JITTER
__device__ __forceinline__ void jitter(uint32_t salt) {
  uint32_t c;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
  c = (c ^ salt ^ (blockIdx.x * 2654435761u)) * 2246822519u;
  if ((c >> 28) < 5) __nanosleep(200 + ((c >> 8) & 4095));
}
#define JITTER(salt) do { if ((MMD_ATTN_JITTER) & ((salt) >> 12)) jitter(salt); } while (0)   // 1 loaders, 2 MMA issuers, 4 softmax (salt's top nibble)
#else
#define JITTER(salt)
#endif

// warpgroup register reallocation: the kernel starts with 128 registers per thread (512 threads); the loader warpgroup
// and the MMA warpgroup hand most of theirs to the two softmax warpgroups (64 scores + 64..128 outputs per thread).
#ifdef MMD_NO_SETMAXNREG   // diagnostic builds only (the softmax warps then spill)
template <int N> __device__ __forceinline__ void reg_dec() {}
template <int N> __device__ __forceinline__ void reg_inc() {}
#else
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
#endif
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// named barriers (ids 1, 2; id 0 is __syncthreads): the two softmax groups hand the MUFU to each other
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Debug timeline (only in builds with -DMMD_ATTN_TRACE, see tools/trace_attn.py): CTA (0,0,0) records %clock at the
// pipeline's hand-over points, [role][tile][event].
#ifdef MMD_ATTN_TRACE
__device__ uint32_t* g_attn_trace = nullptr;
__device__ __forceinline__ uint32_t* trace_base() {
  return (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_attn_trace : nullptr;
}
__device__ __forceinline__ void trace_ev(uint32_t* tr, int role, int j, int e) {
  if (tr != nullptr && j < 64) {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    tr[(role * 64 + j) * 8 + e] = c;
  }
}
#define TRACE_INIT uint32_t* const trace_ptr = trace_base()
#define TRACE_EV(role, j, e) trace_ev(trace_ptr, role, j, e)
#else
#define TRACE_INIT
#define TRACE_EV(role, j, e)
#endif

// Shared-memory layout and mbarrier indices for the two launch shapes:
//   AttnCfg<2, 3>  persistent CTAs walking several work items: Q double-buffered (the loaders fetch the next item's Q while
//                  this one finishes), K and V rings of 3:        2 x 64 KB + 2 x 3 x 16 KB = 224 KB
//   AttnCfg<1, 5>  at most one work item per CTA (short single-frame decoder steps): one Q buffer, rings of 5, so every
//                  tile of a short split is in flight at once:     64 KB + 2 x 5 x 16 KB = 224 KB
template <int QBUF_, int RING_>
struct AttnCfg {
  static constexpr int QBUF = QBUF_, RING = RING_;
  // byte offsets from the 1024-B aligned base
  static constexpr int Q = 0;                                   // [buffer][qi][2 regions]
  static constexpr int QBUF_BYTES = TA_QT * 2 * TA_QREGION;
  static constexpr int K = Q + QBUF * QBUF_BYTES;               // ring of [2 regions]
  static constexpr int V = K + RING * 2 * TA_KREGION;
  static constexpr int BARS = V + RING * 2 * TA_KREGION;
  static constexpr int TOTAL = BARS + 512 + 1024;               // barriers + alignment slack
  // barrier indices
  static constexpr int BAR_Q_FULL = 0, BAR_Q_EMPTY = BAR_Q_FULL + QBUF;   // [buffer]
  static constexpr int BAR_K_FULL = BAR_Q_EMPTY + QBUF, BAR_K_EMPTY = BAR_K_FULL + RING, BAR_V_FULL = BAR_K_EMPTY + RING,
                       BAR_V_EMPTY = BAR_V_FULL + RING;
  static constexpr int BAR_S_FULL = BAR_V_EMPTY + RING;         // [qi][2]
  static constexpr int BAR_P_FULL = BAR_S_FULL + 2 * TA_QT;     // [qi][2]
  // [qi][2], PV_g commits to [g & 1].  The softmax waits on O_FULL only when it rescales and at the end of an item, i.e. NOT
  // for every phase, and a parity wait is only meaningful if the waiter knows the barrier's phase to within one.  With one
  // barrier per parity of g that holds: when PV_k is awaited, PV_{k-2} (the previous phase of the same barrier) is known to
  // be complete (S_{k+1} or S_k has been seen, and it is produced after PV_{k-2}) and PV_{k+2} cannot have been issued.
  // (A single barrier was a latent race: a slow MMA thread could be two phases behind and the wait passed at once.)
  static constexpr int BAR_O_FULL = BAR_P_FULL + 2 * TA_QT;
  static constexpr int BAR_O_EMPTY = BAR_O_FULL + 2 * TA_QT;    // [qi]: the softmax group has read the finished O out of TMEM
  static constexpr int BAR_COUNT = BAR_O_EMPTY + TA_QT;
  static_assert(8 * BAR_COUNT + 8 <= 512, "barrier area");
  static_assert(TOTAL <= 232448, "shared memory per CTA");
};
using AttnCfgPersistent = AttnCfg<2, 3>;
using AttnCfgSingle = AttnCfg<1, 5>;

struct VitAttnParams {
  const __nv_bfloat16* qkv;   // [T*S, 3*H*DH]
  __nv_bfloat16* out;         // [T*S, H*DH] or [T*S, 2*H*DH] (hi | lo)
  int S, H, split_hi_lo;
  float scale_log2e;
};

struct PagedAttnParams {
  const __nv_bfloat16* q;         // [total_q, Hq, 128]
  const __nv_bfloat16* kv_layer;  // [page][2][Hkv][64][128]
  const int* stream_desc;         // [n_streams,4] {q_start, n_q, kv_len, table_off}
  const int* block_tables;
  float* o_part;                  // [n_splits, total_q*Hq, 128]
  float* ml_part;                 // [n_splits, total_q*Hq, 2]
  int Hq, Hkv, n_splits, grid_x;   // grid_x = query blocks x kv heads
  long long part_stride_rows;
  float scale_log2e;
};

// The CTA is persistent: it walks the work items blockIdx.x, blockIdx.x + gridDim.x, ... (an item = 256 query rows of one
// head x its key range).  Every role runs the same item loop with its own running counters, so the loaders are already
// fetching the next item's Q (second Q buffer) and K/V tiles while the softmax groups finish and store the current one,
// and the MMA threads issue the next item's first two QK^T products during that epilogue: the per-item prologue
// (~6k clk until the first S tile) disappears from the critical path after the first item.
template <int DH, typename Front, typename C, typename Params>
__device__ __forceinline__ void attention_pipeline(const Params& prm, int n_items, uint32_t smem_base, uint32_t tmem_base) {
  constexpr int DHP = (DH + 15) / 16 * 16;      // head dim padded to the UMMA K/N granularity
  constexpr int CH = DH / 8;                     // 16-B chunks per row
  // When the head dim has padding columns (72 -> 80), column DH of V holds 1.0 for every key, so O[:, DH] accumulates
  // the row sum of the bf16-ROUNDED probabilities on the tensor core: no FADDs in the softmax, and the sum is rescaled
  // together with O.
  constexpr bool ONES_COL = (DHP > DH);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bars = smem_base + C::BARS;
  auto bar = [&](int i) { return bars + 8u * i; };
  TRACE_INIT;

  if (warp >= TA_SOFTMAX_WARPS && warp < TA_SOFTMAX_WARPS + TA_LOADER_WARPS) {
    // ======================= loaders =======================
    // Two independent streams of 64 threads: warps 0-1 feed the K ring, warps 2-3 the V ring (K tiles are needed two
    // rounds before the matching V tile, so each stream runs as far ahead as its own ring allows); all four load Q first.
    // Work is split in 16-B units with consecutive lanes on consecutive chunks of a row, so one warp instruction reads
    // whole rows (full sectors) instead of one chunk from each of 32 rows.
    // Completion is tracked by the mbarriers themselves (cp.async.mbarrier.arrive.noinc: the arrival fires when this
    // thread's copies so far have landed), so a loader thread never waits for data: it only blocks on a free ring slot
    // or Q buffer and runs ahead of the MMA threads, across item boundaries (a tile takes ~3000 clk to land, one is
    // consumed every ~1000-1800 clk).
    reg_dec<72>();
    const int lt = threadIdx.x - 32 * TA_SOFTMAX_WARPS;      // 0..127
    const bool is_v = lt >= 64;
    const int lrow = lt & 63;
    const int ring = is_v ? C::RING : C::RING;
    const int bar_full = is_v ? C::BAR_V_FULL : C::BAR_K_FULL, bar_empty = is_v ? C::BAR_V_EMPTY : C::BAR_K_EMPTY;
    const uint32_t ring_base = smem_base + (is_v ? C::V : C::K);
    int g = 0;                                               // tiles this stream has loaded so far (all items)
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      Front fe(prm, item);
      const int n_tiles = fe.n_tiles();
      const int qb = it % C::QBUF;
      // the Q buffer is free once both softmax groups have finished the item that used it (their staged stores included)
      mbar_wait(bar(C::BAR_Q_EMPTY + qb), ((it / C::QBUF) & 1) ^ 1);
      const uint32_t qbase = smem_base + C::Q + qb * C::QBUF_BYTES;
#pragma unroll 4
      for (int u = lt; u < TA_QT * TA_BM * CH; u += 32 * TA_LOADER_WARPS) {
        const int r = u / CH, c = u % CH;                    // row within the CTA's 256 query rows, 16-B chunk
        const __nv_bfloat16* src = fe.q_row(r);
        const uint32_t dst = qbase + (r / TA_BM) * 2 * TA_QREGION + (c >> 3) * TA_QREGION + sw128_off(r % TA_BM, c & 7);
        cp_async16(dst, src ? src + c * 8 : fe.any_ptr(), src ? 16 : 0);
      }
      cp_async_mbar_arrive(bar(C::BAR_Q_FULL + qb));
      for (int j = 0; j < n_tiles; ++j, ++g) {
        const int st = g % ring;
        if (lrow == 0) TRACE_EV(4 + is_v, g, 0);
        JITTER(0x1000u | (g & 0xfff));
        mbar_wait(bar(bar_empty + st), ((g / ring) & 1) ^ 1);
        if (lrow == 0) TRACE_EV(4 + is_v, g, 1);
        const uint32_t dstb = ring_base + st * 2 * TA_KREGION;
        int valid;                                           // rows past `valid` are zero-filled
        const __nv_bfloat16* tile = fe.kv_tile(j, is_v, valid);
        const long long rstride = fe.kv_row_stride();
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int u = i * 64 + lrow, r = u / CH, c = u % CH;
          const bool ok = r < valid;
          cp_async16(dstb + (c >> 3) * TA_KREGION + sw128_off(r, c & 7), ok ? tile + r * rstride + c * 8 : fe.any_ptr(), ok ? 16 : 0);
        }
        cp_async_mbar_arrive(bar(bar_full + st));
        if (lrow == 0) TRACE_EV(4 + is_v, g, 2);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();   // nothing may still be writing this CTA's shared memory when it exits
  } else if (warp >= TA_SOFTMAX_WARPS + TA_LOADER_WARPS) {
    // ======================= MMA issuers: one thread per query tile (warps 12, 13; warps 14, 15 only give up their
    // registers).  A wait on an mbarrier costs ~200 clk even when it has long completed, so one thread serving both
    // tiles (6 waits + 24 MMAs per key tile) was slower than the tensor pipe; each thread's own MMAs stay in order,
    // which is all the S/P buffer reuse relies on. =======================
    reg_dec<56>();
    const int qi = warp - (TA_SOFTMAX_WARPS + TA_LOADER_WARPS);
    if (qi < TA_QT && elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(TA_BM, TA_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_mn_b(TA_BM, DHP, MMD_EXP_F16X2 ? 0u : 1u);
      // One thread issues everything, so its instruction count per tile is on the critical path: the descriptors' high
      // words are constants and the low words (start address | LBO) advance by immediates.
      constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);              // SBO | version 1 | SWIZZLE_128B
      auto desc = [](uint32_t lo) { return (static_cast<uint64_t>(HI) << 32) | lo; };
      auto lo_k = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };                       // K-major
      auto lo_mn = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | ((uint32_t)(TA_KREGION >> 4) << 16); };   // MN-major
      const uint32_t q_lo = lo_k(smem_base + C::Q), k_lo = lo_k(smem_base + C::K), v_lo = lo_mn(smem_base + C::V);
      int g0 = 0;            // key tiles of all earlier items: S/P/O barrier phases and the K/V ring positions run on
      int busy_items = 0;    // items that had at least one tile (O_EMPTY phases)
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        Front fe(prm, item);
        const int n_tiles = fe.n_tiles();
        const int qb = it % C::QBUF;
        const uint32_t qa = q_lo + ((qb * C::QBUF_BYTES + qi * 2 * TA_QREGION) >> 4);
        auto issue_qk = [&](int j) {
          const int g = g0 + j, st = g % C::RING, sb = g & 1;
          TRACE_EV(2 + qi, g, 4);
          mbar_wait(bar(C::BAR_K_FULL + st), (g / C::RING) & 1);
          fence_proxy_async_smem();                          // cp.async (generic proxy) data -> tcgen05.mma operand reads
          tc_fence_after();
          TRACE_EV(2 + qi, g, 5);
          // No "S buffer free" barrier: QK_g is issued after PV_{g-2} (which read P_{g-2} from this buffer, itself written
          // after S_{g-2} had been read), and the tensor pipe executes this thread's MMAs in issue order.
          const uint32_t ka = k_lo + st * (2 * TA_KREGION >> 4);
          const uint32_t d = tmem_base + qi * TA_TMEM_PER_Q + sb * TA_BN;
#pragma unroll
          for (int k = 0; k < DHP / 16; ++k)
            umma_f16(d, desc(qa + (k >> 2) * (TA_QREGION >> 4) + 2 * (k & 3)), desc(ka + (k >> 2) * (TA_KREGION >> 4) + 2 * (k & 3)), idesc_qk,
                     k > 0 ? 1u : 0u);
          umma_commit(bar(C::BAR_S_FULL + qi * 2 + sb));
          umma_commit(bar(C::BAR_K_EMPTY + st));                // the slot is free once both tiles' issuers have arrived
          TRACE_EV(2 + qi, g, 6);
        };
        mbar_wait(bar(C::BAR_Q_FULL + qb), (it / C::QBUF) & 1);
        fence_proxy_async_smem();
        tc_fence_after();
        for (int j = 0; j < 2 && j < n_tiles; ++j) issue_qk(j);
        for (int j = 0; j < n_tiles; ++j) {
          const int g = g0 + j, st = g % C::RING;
          JITTER(0x2000u | (g & 0xfff));
          TRACE_EV(2 + qi, g, 0);
          mbar_wait(bar(C::BAR_V_FULL + st), (g / C::RING) & 1);   // usually long complete: take its latency before P arrives
          TRACE_EV(2 + qi, g, 1);
          mbar_wait(bar(C::BAR_P_FULL + qi * 2 + (g & 1)), (g >> 1) & 1);
          if (j == 0) {     // the previous item's O must have been read out of TMEM before this PV overwrites it
            mbar_wait(bar(C::BAR_O_EMPTY + qi), (busy_items & 1) ^ 1);
            ++busy_items;
          }
          TRACE_EV(2 + qi, g, 2);
          fence_proxy_async_smem();
          tc_fence_after();
          // P_g (bf16) sits in TMEM over the first 32 columns of the S buffer it was computed from: 8 columns per K=16 step
          const uint32_t tP = tmem_base + qi * TA_TMEM_PER_Q + (g & 1) * TA_BN;
          const uint32_t va = v_lo + st * (2 * TA_KREGION >> 4);
#pragma unroll
          for (int k = 0; k < TA_BN / 16; ++k)
            umma_f16_ts(tmem_base + qi * TA_TMEM_PER_Q + 2 * TA_BN, tP + 8 * k, desc(va + k * (2048 >> 4)), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(bar(C::BAR_O_FULL + qi * 2 + (g & 1)));
          umma_commit(bar(C::BAR_V_EMPTY + st));
          TRACE_EV(2 + qi, g, 3);
          if (j + 2 < n_tiles) issue_qk(j + 2);
        }
        g0 += n_tiles;
      }
    }
    __syncwarp();
  } else {
    // ======================= softmax: one thread per query row, one warpgroup per query tile =======================
    reg_inc<192>();
    const int qi = warp >> 2;
    const int row = (warp & 3) * 32 + lane;                 // row within the query tile = TMEM lane
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + qi * TA_TMEM_PER_Q;
    // MUFU.EX2 (16 lanes/clk/SM) is the slowest pipe of this kernel: one key tile costs 2 x 128 x 64 exponentials =
    // 1024 clk of it, as much as the tile's tensor work.  A warp issues in order, so while its MUFU instructions queue
    // nothing else of that warp moves; the second group's TMEM loads / max / barrier traffic fill those slots.
    struct Tile { uint32_t a[32], b[32]; };
    auto val = [](const Tile& t, int i) { return __uint_as_float(i < 32 ? t.a[i] : t.b[i - 32]); };
    // The two groups take turns in the exponential phase (named barriers 1 + qi: "group qi may go"): a group alone
    // streams its MUFU work almost back to back, and the other group's non-MUFU phases run underneath instead of both
    // groups queueing on the MUFU and then both leaving it idle.  The token keeps circulating across items.
    constexpr int NSM = 32 * TA_SOFTMAX_WARPS;
    int total_tiles = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) total_tiles += Front(prm, item).n_tiles();
#ifndef MMD_ATTN_NO_ALT
    if (qi == 1 && total_tiles > 0) named_bar_arrive(1, NSM);
#endif
    int g0 = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      Front fe(prm, item);
      const int n_tiles = fe.n_tiles();
      const int key_lim = fe.key_limit(qi * TA_BM + row);  // tile-local key indices > key_lim are masked for this row
      const float sl2 = fe.scale_log2e();
      float m_ref = -INFINITY, l_run = 0.f;
      Tile cur;
      for (int j = 0; j < n_tiles; ++j) {
        const int g = g0 + j, sb = g & 1;
        if (row == 0) TRACE_EV(qi, g, 0);
        JITTER(0x4000u | ((g + (warp << 6)) & 0xfff));
        mbar_wait(bar(C::BAR_S_FULL + qi * 2 + sb), (g >> 1) & 1);     // S_g: wait, then read the row into registers
        tc_fence_after();
        tmem_ld_32x32b_x32(t_row + sb * TA_BN, cur.a);
        tmem_ld_32x32b_x32(t_row + sb * TA_BN + 32, cur.b);
        if (row == 0) TRACE_EV(qi, g, 1);
        tmem_ld_wait();
        {
          // masking only in the tiles that reach beyond some row's limit (warp-uniform test; ISETP/SEL per element only there)
          const int k0 = j * TA_BN;
          if (__any_sync(0xffffffffu, k0 + TA_BN - 1 > key_lim)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              cur.a[i] = (k0 + i > key_lim) ? 0xff800000u : cur.a[i];
              cur.b[i] = (k0 + 32 + i > key_lim) ? 0xff800000u : cur.b[i];
            }
          }
        }
        if (row == 0) TRACE_EV(qi, g, 2);
        float mx;
        {                                                    // four chains of 3-input FMNMX3
          float m0 = fmaxf(val(cur, 0), val(cur, 1)), m1 = fmaxf(val(cur, 2), val(cur, 3)), m2 = fmaxf(val(cur, 4), val(cur, 5)),
                m3 = fmaxf(val(cur, 6), val(cur, 7));
#pragma unroll
          for (int i = 8; i < TA_BN; i += 8) {
            m0 = fmaxf(m0, fmaxf(val(cur, i), val(cur, i + 1)));
            m1 = fmaxf(m1, fmaxf(val(cur, i + 2), val(cur, i + 3)));
            m2 = fmaxf(m2, fmaxf(val(cur, i + 4), val(cur, i + 5)));
            m3 = fmaxf(m3, fmaxf(val(cur, i + 6), val(cur, i + 7)));
          }
          mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        }
        const float m_new = fmaxf(m_ref, mx);
        const bool need = j > 0 && m_new > m_ref && (m_ref == -INFINITY || (m_new - m_ref) * sl2 > 8.0f);
        if (j == 0) {
          m_ref = m_new;
        } else if (__any_sync(0xffffffffu, need)) {
          // lazy rescaling: O (TMEM) and l move to the new exponent base.  tcgen05.ld/st are .sync.aligned, so the whole
          // warp takes this path together; rows that do not need it rescale by 1.  PV_{g-1} must have completed first.
          mbar_wait(bar(C::BAR_O_FULL + qi * 2 + ((g - 1) & 1)), ((g - 1) >> 1) & 1);
          tc_fence_after();
          const float f = !need ? 1.f : ((m_ref == -INFINITY) ? 0.f : exp2f((m_ref - m_new) * sl2));
#pragma unroll
          for (int c0 = 0; c0 < DHP; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_row + 2 * TA_BN + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st_32x32b_x16(t_row + 2 * TA_BN + c0, v);
          }
          tmem_st_wait();
          if (need) {
            l_run *= f;
            m_ref = m_new;
          }
        }
#ifndef MMD_ATTN_NO_ALT
        named_bar_sync(1 + qi, NSM);
#endif
        if (row == 0) TRACE_EV(qi, g, 3);
        const float msc = (m_ref == -INFINITY) ? 0.f : m_ref * sl2;
        // probabilities -> bf16 P tile, written back into TMEM over the S buffer they came from (two keys per 32-bit
        // column).  PV_{g-2}, the last reader of these columns, completed before S_g was produced: no wait needed, and
        // the softmax of the next tile can start while PV_g is still running.
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < TA_BN / 32; ++c) {         // 16 columns = 32 keys per store
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = fmaf(val(cur, 32 * c + i), sl2, -msc);
          uint32_t pk[16];
#if MMD_EXP_F16X2
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = ex2_f16x2(x[2 * i], x[2 * i + 1]);
          if constexpr (!ONES_COL) {
            // row sum of the ROUNDED probabilities: pairwise f16x2 tree over the 32 keys (depth 4), then fp32
            uint32_t t8[8], t4[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = hadd2_u32(pk[2 * i], pk[2 * i + 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) t4[i] = hadd2_u32(t8[2 * i], t8[2 * i + 1]);
            rs += h2_sum_f32(hadd2_u32(hadd2_u32(t4[0], t4[1]), hadd2_u32(t4[2], t4[3])));
          }
#else
#pragma unroll
          for (int i = 0; i < 16; ++i) {
#if MMD_EXP_MODE == 3   // timing experiment only: no exponential at all
            const float p0 = x[2 * i], p1 = x[2 * i + 1];
#else
            const float p0 = exp_on_fma_pipe(i) ? exp2_poly(x[2 * i]) : exp2f(x[2 * i]);
            const float p1 = exp_on_fma_pipe(i) ? exp2_poly(x[2 * i + 1]) : exp2f(x[2 * i + 1]);
#endif
#ifdef MMD_P_TRUNC      // experiment: bf16 pair by taking the high halves (one PRMT on the ALU pipe) instead of F2FP.BF16.PACK_AB
            pk[i] = __byte_perm(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
#else
            __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
            pk[i] = *reinterpret_cast<uint32_t*>(&h);
#endif
            if constexpr (!ONES_COL) {
              // row sum of the ROUNDED probabilities (numerator and denominator stay consistent under the stale base)
              rs += __uint_as_float(pk[i] << 16) + __uint_as_float(pk[i] & 0xffff0000u);
            }
          }
#endif
          tmem_st_32x32b_x16(t_row + sb * TA_BN + 16 * c, pk);
        }
        l_run += rs;
#ifndef MMD_ATTN_NO_ALT
        if (qi == 0 || g + 1 < total_tiles) named_bar_arrive(2 - qi, NSM);
#endif
        if (row == 0) TRACE_EV(qi, g, 4);
        tmem_st_wait();
        if (row == 0) TRACE_EV(qi, g, 5);
        tc_fence_before();
        // one P_FULL barrier per S buffer: tile g+2's arrivals cannot start before the MMA thread has consumed tile g's
        // (S_{g+2} is produced after PV_g), so a phase can never be skipped and no wait for PV_{g-1} is needed here
        mbar_arrive(bar(C::BAR_P_FULL + qi * 2 + sb));
        if (row == 0) TRACE_EV(qi, g, 6);
      }
      if (row == 0) TRACE_EV(6, qi + 2 * (it & 3), 2);
      float o[DHP];
      if (n_tiles > 0) {
        mbar_wait(bar(C::BAR_O_FULL + qi * 2 + ((g0 + n_tiles - 1) & 1)), ((g0 + n_tiles - 1) >> 1) & 1);
        tc_fence_after();
        {   // all loads in flight, one wait
          uint32_t v[DHP / 16][16];
#pragma unroll
          for (int c = 0; c < DHP / 16; ++c) tmem_ld_32x32b_x16(t_row + 2 * TA_BN + 16 * c, v[c]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < DHP / 16; ++c)
#pragma unroll
            for (int i = 0; i < 16; ++i) o[16 * c + i] = __uint_as_float(v[c][i]);
        }
        if constexpr (ONES_COL) l_run = o[DH];
        tc_fence_before();
        mbar_arrive(bar(C::BAR_O_EMPTY + qi));                  // the next item's first PV may overwrite O now
      } else {
#pragma unroll
        for (int i = 0; i < DHP; ++i) o[i] = 0.f;
      }
      if (row == 0) TRACE_EV(6, qi + 2 * (it & 3), 3);
      // Output goes through the group's own (now idle) part of this item's Q buffer so that global stores are whole rows
      // per warp instruction instead of one 16-B piece from each of 32 rows; named barrier 3 + qi synchronises the group.
      fe.store(qi, row, o, m_ref, l_run, smem_base + C::Q + (it % C::QBUF) * C::QBUF_BYTES + qi * 2 * TA_QREGION, 3 + qi);
      mbar_arrive(bar(C::BAR_Q_EMPTY + it % C::QBUF));           // this thread is done with the staging area: once all 256
                                                              // softmax threads are, the loaders may refill the Q buffer
      if (row == 0) TRACE_EV(6, qi + 2 * (it & 3), 4);
      g0 += n_tiles;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// front end: SigLIP (packed qkv, non-causal)
// ---------------------------------------------------------------------------------------------------------------
template <int DH>
struct VitFront {
  const VitAttnParams& p;
  int t, h, q0;
  const __nv_bfloat16 *gQ, *gK, *gV;
  int row_stride;
  __device__ VitFront(const VitAttnParams& p_, int item) : p(p_) {
    const int qblocks = (p.S + TA_BM * TA_QT - 1) / (TA_BM * TA_QT);   // item = (frame, head, query block), query block fastest
    q0 = (item % qblocks) * TA_BM * TA_QT; h = (item / qblocks) % p.H; t = item / (qblocks * p.H);
    row_stride = 3 * p.H * DH;
    const __nv_bfloat16* base = p.qkv + (long long)t * p.S * row_stride;
    gQ = base + h * DH; gK = base + (p.H + h) * DH; gV = base + (2 * p.H + h) * DH;
  }
  __device__ int n_tiles() const { return (p.S + TA_BN - 1) / TA_BN; }
  __device__ const __nv_bfloat16* any_ptr() const { return p.qkv; }
  __device__ const __nv_bfloat16* q_row(int r) const { return q0 + r < p.S ? gQ + (long long)(q0 + r) * row_stride : nullptr; }
  __device__ const __nv_bfloat16* kv_tile(int j, bool is_v, int& valid) const {
    valid = min(TA_BN, p.S - j * TA_BN);
    return (is_v ? gV : gK) + (long long)j * TA_BN * row_stride;
  }
  __device__ long long kv_row_stride() const { return row_stride; }
  __device__ int key_limit(int) const { return p.S - 1; }
  __device__ float scale_log2e() const { return p.scale_log2e; }
  template <int DHP>
  __device__ void store(int qi, int row, const float (&o)[DHP], float, float l, uint32_t stage, int bar_id) const {
    constexpr int CHO = DH / 8, PITCH = DH * 2;              // 16-B chunks and bytes per staged row
    const float inv = 1.f / l;
    const int out_stride = (p.split_hi_lo ? 2 : 1) * p.H * DH;
    const int halves = p.split_hi_lo ? 2 : 1;
    for (int half = 0; half < halves; ++half) {
#pragma unroll
      for (int c = 0; c < CHO; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = o[c * 8 + 2 * i] * inv, b = o[c * 8 + 2 * i + 1] * inv;
          __nv_bfloat162 hv = __floats2bfloat162_rn(a, b);
          if (half == 1) {                                   // lo = what the bf16 rounding of hi dropped
            const float2 hf = __bfloat1622float2(hv);
            hv = __floats2bfloat162_rn(a - hf.x, b - hf.y);
          }
          w[i] = *reinterpret_cast<uint32_t*>(&hv);
        }
        sts128(stage + row * PITCH + c * 16, w[0], w[1], w[2], w[3]);
      }
      named_bar_sync(bar_id, TA_BM);
      __nv_bfloat16* ob = p.out + (long long)t * p.S * out_stride + half * p.H * DH + h * DH;
      for (int u = row; u < TA_BM * CHO; u += TA_BM) {
        const int r = u / CHO, c = u % CHO, qr = q0 + qi * TA_BM + r;
        if (qr < p.S) *reinterpret_cast<uint4*>(ob + (long long)qr * out_stride + c * 8) = lds128(stage + r * PITCH + c * 16);
      }
      if (half + 1 < halves) named_bar_sync(bar_id, TA_BM);  // staging is reused for the lo half
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// front end: Qwen2 decoder over the paged KV pool (GQA rows stacked, causal, split-KV)
// ---------------------------------------------------------------------------------------------------------------
struct PagedFront {
  static constexpr int DH = 128, PAGE = 64;
  const PagedAttnParams& p;
  int G, kvh, sp, q_start, n_q, kv_len, R, r_base, past, t_begin, t_end;
  const int* table;
  __device__ PagedFront(const PagedAttnParams& p_, int item) : p(p_) {
    // item = (stream, split, query block x kv head), the last fastest
    const int bx = item % p.grid_x;
    G = p.Hq / p.Hkv;
    kvh = bx % p.Hkv;
    const int qt = bx / p.Hkv;
    sp = (item / p.grid_x) % p.n_splits;
    const int st = item / (p.grid_x * p.n_splits);
    q_start = p.stream_desc[st * 4 + 0]; n_q = p.stream_desc[st * 4 + 1]; kv_len = p.stream_desc[st * 4 + 2];
    table = p.block_tables + p.stream_desc[st * 4 + 3];
    R = n_q * G;
    r_base = qt * TA_BM * TA_QT;
    past = kv_len - n_q;
    int vis = 0;
    if (r_base < R) {
      const int last_row = min(R, r_base + TA_BM * TA_QT) - 1;
      vis = min((kv_len + TA_BN - 1) / TA_BN, (past + last_row / G) / TA_BN + 1);
    }
    const int per = (vis + p.n_splits - 1) / p.n_splits;
    t_begin = sp * per;
    t_end = min(vis, t_begin + per);
    if (t_end < t_begin) t_end = t_begin;
  }
  __device__ int n_tiles() const { return t_end - t_begin; }
  __device__ const __nv_bfloat16* any_ptr() const { return p.q; }
  // stacked row rr = token * G + head-in-group  ->  row of the [total_q * Hq] query / output matrices
  __device__ long long grow_of(int rr) const {
    const int tok = (G == 7) ? rr / 7 : rr / G;         // Qwen2-7B: 28 query heads on 4 KV heads (constant division)
    return (long long)(q_start + tok) * p.Hq + kvh * G + (rr - tok * G);
  }
  __device__ const __nv_bfloat16* q_row(int r) const {
    const int rr = r_base + r;
    return rr < R ? p.q + grow_of(rr) * DH : nullptr;
  }
  __device__ const __nv_bfloat16* kv_tile(int j, bool is_v, int& valid) const {
    const int tile = t_begin + j;                       // one key tile == one 64-token page
    valid = min(TA_BN, kv_len - tile * TA_BN);          // pool memory past the end may hold anything: zero-fill
    return p.kv_layer + (((long long)table[tile] * 2 + (is_v ? 1 : 0)) * p.Hkv + kvh) * PAGE * DH;
  }
  __device__ long long kv_row_stride() const { return DH; }
  // masking works on tile-local key indices j*TA_BN + c: shift the causal limit into that frame
  __device__ int key_limit(int row) const { return past + min(r_base + row, max(R - 1, 0)) / G - t_begin * TA_BN; }
  __device__ float scale_log2e() const { return p.scale_log2e; }
  template <int DHP>
  __device__ void store(int qi, int row, const float (&o)[DHP], float m, float l, uint32_t stage, int bar_id) const {
    {
      const int rr = r_base + qi * TA_BM + row;
      if (rr < R) {
        float* mp = p.ml_part + ((long long)sp * p.part_stride_rows + grow_of(rr)) * 2;
        mp[0] = (m == -INFINITY) ? -INFINITY : m * p.scale_log2e;
        mp[1] = l;
      }
    }
    // fp32 partial rows, 64 columns per pass (the Q region holds 128 x 256 B); chunks are XOR-swizzled by row so that
    // both the row-per-thread writes and the chunk-per-lane reads are conflict-free
    float* ob = p.o_part + (long long)sp * p.part_stride_rows * DH;
#pragma unroll
    for (int pass = 0; pass < DH / 64; ++pass) {
#pragma unroll
      for (int c = 0; c < 16; ++c)
        sts128(stage + row * 256 + ((c ^ (row & 15)) << 4), __float_as_uint(o[pass * 64 + 4 * c]), __float_as_uint(o[pass * 64 + 4 * c + 1]),
               __float_as_uint(o[pass * 64 + 4 * c + 2]), __float_as_uint(o[pass * 64 + 4 * c + 3]));
      named_bar_sync(bar_id, TA_BM);
#pragma unroll 4
      for (int u = row; u < TA_BM * 16; u += TA_BM) {
        const int r = u >> 4, c = u & 15, rr = r_base + qi * TA_BM + r;
        if (rr < R) *reinterpret_cast<uint4*>(ob + grow_of(rr) * DH + pass * 64 + c * 4) = lds128(stage + r * 256 + ((c ^ (r & 15)) << 4));
      }
      if (pass + 1 < DH / 64) named_bar_sync(bar_id, TA_BM);
    }
  }
};

template <int DH, typename Params, typename Front, typename C>
__global__ void __launch_bounds__(TA_THREADS, 1) attn_tcgen05_kernel(const Params p, const int n_items) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + C::BARS;
  const uint32_t tmem_slot = bars + 8u * C::BAR_COUNT;
  const int warp = threadIdx.x >> 5;
  TRACE_INIT;
  griddep_launch_dependents();
  if (threadIdx.x == 0) TRACE_EV(6, 0, 0);
  if (threadIdx.x == 0) {
    constexpr uint32_t NLOAD = 32 * TA_LOADER_WARPS / 2;   // one 64-thread stream per ring
    for (int i = 0; i < C::QBUF; ++i) {
      mbar_init(bars + 8u * (C::BAR_Q_FULL + i), 32 * TA_LOADER_WARPS);
      mbar_init(bars + 8u * (C::BAR_Q_EMPTY + i), 32 * TA_SOFTMAX_WARPS);
    }
    for (int s = 0; s < C::RING; ++s) {
      mbar_init(bars + 8u * (C::BAR_K_FULL + s), NLOAD);
      mbar_init(bars + 8u * (C::BAR_K_EMPTY + s), TA_QT);
    }
    for (int s = 0; s < C::RING; ++s) {
      mbar_init(bars + 8u * (C::BAR_V_FULL + s), NLOAD);
      mbar_init(bars + 8u * (C::BAR_V_EMPTY + s), TA_QT);
    }
    for (int i = 0; i < 2 * TA_QT; ++i) {
      mbar_init(bars + 8u * (C::BAR_S_FULL + i), 1);
      mbar_init(bars + 8u * (C::BAR_P_FULL + i), 128);
    }
    for (int i = 0; i < 2 * TA_QT; ++i) mbar_init(bars + 8u * (C::BAR_O_FULL + i), 1);
    for (int i = 0; i < TA_QT; ++i) mbar_init(bars + 8u * (C::BAR_O_EMPTY + i), TA_BM);
    fence_mbar_init();
  }
  // Head-dim padding (72 -> 80): the loaders never write those chunks, so they are set once here: zero everywhere, and
  // 1.0 in column DH of V in every ring stage (the "ones column", see attention_pipeline).
  if constexpr (((DH + 15) / 16 * 16) > DH) {
    static_assert(DH % 8 == 0 && ((DH + 15) / 16 * 16) - DH == 8, "one 16-B padding chunk");
    uint8_t* sm = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr int PC = DH / 8;                                           // index of the padding chunk
    constexpr int QROWS = C::QBUF * TA_QT * TA_BM;           // both Q buffers (tiles are contiguous: [buffer][qi])
    constexpr int ROWS = QROWS + (C::RING + C::RING) * TA_BN;
    for (int i = threadIdx.x; i < ROWS; i += TA_THREADS) {
      uint32_t off;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (i < QROWS) {
        off = C::Q + (i / TA_BM) * 2 * TA_QREGION + (PC >> 3) * TA_QREGION + sw128_off(i % TA_BM, PC & 7);
      } else if (i < QROWS + C::RING * TA_BN) {
        const int k = i - QROWS;
        off = C::K + (k / TA_BN) * 2 * TA_KREGION + (PC >> 3) * TA_KREGION + sw128_off(k % TA_BN, PC & 7);
      } else {
        const int k = i - QROWS - C::RING * TA_BN;
        off = C::V + (k / TA_BN) * 2 * TA_KREGION + (PC >> 3) * TA_KREGION + sw128_off(k % TA_BN, PC & 7);
        val.x = 0x3f80u;                                                 // bf16 1.0 in element 0 of the chunk
      }
      *reinterpret_cast<uint4*>(sm + off) = val;
    }
  }
  if (warp == TA_SOFTMAX_WARPS + TA_LOADER_WARPS) tmem_alloc<512>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();
  if (threadIdx.x == 0) TRACE_EV(6, 0, 1);
  attention_pipeline<DH, Front, C>(p, n_items, smem_base, tmem_base);
  tc_fence_before();
  __syncthreads();
  if (warp == TA_SOFTMAX_WARPS + TA_LOADER_WARPS) tmem_dealloc<512>(tmem_base);
  if (threadIdx.x == 0) TRACE_EV(6, 0, 5);
}

}  // namespace

#ifdef MMD_ATTN_TRACE
extern "C" __attribute__((visibility("default"))) int mmd_debug_attn_trace(void* dev_buf) {
  return cudaMemcpyToSymbol(g_attn_trace, &dev_buf, sizeof(void*)) == cudaSuccess ? 0 : -1;
}
#endif

// persistent grid: one CTA per SM (the kernel needs all of an SM's TMEM and shared memory)
static int attn_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <typename Kern, typename Params>
static int launch_attn(Kern kern, bool* attr_done, int smem, const Params& p, int n_items, cudaStream_t s) {
  if (!*attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -4;
    *attr_done = true;
  }
  dim3 grid(min(n_items, attn_num_sms()));
  if (launch_k(kern, grid, dim3(TA_THREADS), smem, s, p, n_items) != cudaSuccess) return -5;
  return 0;
}

// more items than SMs -> persistent shape (double-buffered Q); otherwise every CTA has one item and deeper K/V rings pay more
template <int DH, typename Params, typename Front>
static int launch_attn_auto(const Params& p, int n_items, cudaStream_t s) {
  static PerDeviceFlag attr_p, attr_s;
  if (n_items > attn_num_sms())
    return launch_attn(attn_tcgen05_kernel<DH, Params, Front, AttnCfgPersistent>, &attr_p.cur(), AttnCfgPersistent::TOTAL, p, n_items, s);
  return launch_attn(attn_tcgen05_kernel<DH, Params, Front, AttnCfgSingle>, &attr_s.cur(), AttnCfgSingle::TOTAL, p, n_items, s);
}

int launch_vit_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int T, int S, int H, int dh, int split_hi_lo, cudaStream_t s) {
  if (T <= 0) return 0;
  if (dh != 72) return -2;
  VitAttnParams p;
  p.qkv = qkv; p.out = out; p.S = S; p.H = H; p.split_hi_lo = split_hi_lo;
  p.scale_log2e = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  const int n_items = ((S + TA_BM * TA_QT - 1) / (TA_BM * TA_QT)) * H * T;
  return launch_attn_auto<72, VitAttnParams, VitFront<72>>(p, n_items, s);
}

int launch_kv_attention_tc_main(const __nv_bfloat16* q, const __nv_bfloat16* kv_layer, const int* stream_desc, const int* block_tables,
                                int n_streams, int max_n_q, int total_q, float* o_part, float* ml_part, int Hq, int Hkv, int n_splits,
                                cudaStream_t s) {
  PagedAttnParams p;
  p.q = q; p.kv_layer = kv_layer; p.stream_desc = stream_desc; p.block_tables = block_tables; p.o_part = o_part; p.ml_part = ml_part;
  p.Hq = Hq; p.Hkv = Hkv; p.n_splits = n_splits; p.part_stride_rows = (long long)total_q * Hq;
  p.scale_log2e = (1.0f / sqrtf(128.f)) * 1.4426950408889634f;
  const int G = Hq / Hkv;
  p.grid_x = ((max_n_q * G + TA_BM * TA_QT - 1) / (TA_BM * TA_QT)) * Hkv;
  return launch_attn_auto<128, PagedAttnParams, PagedFront>(p, p.grid_x * n_splits * n_streams, s);
}

int kv_attention_tc_q_rows_per_cta() { return TA_BM * TA_QT; }

}  // namespace mmd
