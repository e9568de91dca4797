// Flash attention on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM) for sm_100a.
//
//   S = Q K^T   : UMMA 128 x 128 x 16, A = Q tile (smem, K-major), B = K tile (smem, K-major)      -> TMEM (double-buffered)
//   O_j = P V   : UMMA 128 x DH  x 16, A = P tile (smem, K-major), B = V tile (smem, MN-major: the
//                 same [key][d] image the loader writes for K, no transpose)                        -> TMEM
//   softmax     : 4 warps, ONE THREAD PER QUERY ROW (TMEM lane): tcgen05.ld the row of S, running max / sum in
//                 registers with no shuffles, P written to shared memory as bf16 in the 128-B-swizzled K-major layout,
//                 O accumulated in fp32 registers (O_run = O_run * corr + O_j read back from TMEM).
//   loaders     : 4 warps, 16-B cp.async into the swizzled operand layout (zero fill for rows past the end and for the
//                 head-dim padding 72 -> 80), handed to the async proxy with fence.proxy.async + mbarrier.
//   MMA         : one thread issues every tcgen05.mma and the tcgen05.commit that signals the mbarriers.
//
// Two front ends share the pipeline:
//   * vit:   non-causal attention over the 729 patch tokens of a frame, packed qkv [T*S, 3*H*72]
//            (TF:models/siglip/modeling_siglip.py:252-330), optional hi+lo output for the out-projection;
//   * paged: KV-append attention of the Qwen2 decoder over the paged KV pool with GQA row stacking, bottom-right
//            causal mask and split-KV partial outputs (TF:models/qwen2/modeling_qwen2.py:187-246).
#include "kernels.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <math.h>

namespace mmd {

namespace {

constexpr int TA_BM = 128;        // query rows per CTA = TMEM lanes
constexpr int TA_BN = 128;        // keys per tile
constexpr int TA_REGION = TA_BM * 128;   // bytes of one 64-column (128-B wide) swizzled operand region with 128 rows
constexpr int TA_SOFTMAX_WARPS = 4, TA_LOADER_WARPS = 4;
constexpr int TA_THREADS = 32 * (TA_SOFTMAX_WARPS + TA_LOADER_WARPS + 1);

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// byte offset of 16-B chunk `c` (0..7) of row `r` inside a 128-B-swizzled K-major region (rows 128 B apart, 8-row
// groups 1024 B apart): exactly the image TMA writes with CU_TENSOR_MAP_SWIZZLE_128B.
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// smem descriptor, MN-major operand (here: V as [key][d], d contiguous), 128-B swizzle: LBO = distance between the
// 64-element blocks along N (d), SBO = distance between 8-row groups along K (keys) = 1024 B.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn_b(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

struct AttnSmem {  // offsets from the 1024-B aligned base
  static constexpr int Q = 0;                       // 2 regions
  static constexpr int P = Q + 2 * TA_REGION;       // 2 regions (keys 0-63 / 64-127)
  static constexpr int K0 = P + 2 * TA_REGION;      // per stage: K (2 regions) + V (2 regions)
  static constexpr int STAGE = 4 * TA_REGION;
  static constexpr int BARS = K0 + 2 * STAGE;
  static constexpr int TOTAL = BARS + 256 + 1024;   // + alignment slack
};

enum { BAR_Q_FULL = 0, BAR_KV_FULL = 1, BAR_KV_EMPTY = 3, BAR_S_FULL = 5, BAR_S_EMPTY = 7, BAR_P_FULL = 9, BAR_P_EMPTY = 10,
       BAR_O_FULL = 11, BAR_O_EMPTY = 12, BAR_COUNT = 13 };

// ---- problem descriptions (one per CTA) -------------------------------------------------------------------------
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
  int Hq, Hkv, n_splits;
  long long part_stride_rows;
  float scale_log2e;
};

// The tile loop shared by both front ends.  `Front` provides: n_tiles, loading of Q / K / V rows, the key limit of
// every query row (masking) and the output stage.
template <int DH, typename Front>
__device__ __forceinline__ void attention_pipeline(Front& fe, uint32_t smem_base, uint32_t tmem_base) {
  constexpr int DHP = (DH + 15) / 16 * 16;      // head dim padded to the UMMA K/N granularity
  constexpr int CH = DH / 8;                     // 16-B chunks per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bars = smem_base + AttnSmem::BARS;
  auto bar = [&](int i) { return bars + 8u * i; };
  const int n_tiles = fe.n_tiles();

  if (warp >= TA_SOFTMAX_WARPS && warp < TA_SOFTMAX_WARPS + TA_LOADER_WARPS) {
    // ======================= loaders =======================
    const int lt = threadIdx.x - 32 * TA_SOFTMAX_WARPS;      // 0..127
    constexpr int NL = 32 * TA_LOADER_WARPS;
    // Q: 128 rows x CH chunks
    for (int i = lt; i < TA_BM * CH; i += NL) {
      const int r = i / CH, c = i % CH;
      const __nv_bfloat16* src = fe.q_row(r);
      cp_async16(smem_base + AttnSmem::Q + (c >> 3) * TA_REGION + sw128_off(r, c & 7), src ? src + c * 8 : fe.any_ptr(), src ? 16 : 0);
    }
    cp_async_wait_all();
    fence_proxy_async_smem();
    mbar_arrive(bar(BAR_Q_FULL));
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      mbar_wait(bar(BAR_KV_EMPTY + st), ((j >> 1) & 1) ^ 1);
      const uint32_t kb = smem_base + AttnSmem::K0 + st * AttnSmem::STAGE, vb = kb + 2 * TA_REGION;
      for (int i = lt; i < TA_BN * CH; i += NL) {
        const int r = i / CH, c = i % CH;
        const __nv_bfloat16* ks = fe.k_row(j, r);
        const __nv_bfloat16* vs = fe.v_row(j, r);
        const uint32_t off = (c >> 3) * TA_REGION + sw128_off(r, c & 7);
        cp_async16(kb + off, ks ? ks + c * 8 : fe.any_ptr(), ks ? 16 : 0);
        cp_async16(vb + off, vs ? vs + c * 8 : fe.any_ptr(), vs ? 16 : 0);
      }
      cp_async_wait_all();
      fence_proxy_async_smem();
      mbar_arrive(bar(BAR_KV_FULL + st));
    }
  } else if (warp == TA_SOFTMAX_WARPS + TA_LOADER_WARPS) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(TA_BM, TA_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_mn_b(TA_BM, DHP);
      const uint32_t sQ = smem_base + AttnSmem::Q, sP = smem_base + AttnSmem::P;
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(bar(BAR_KV_FULL + st), (j >> 1) & 1);
        mbar_wait(bar(BAR_S_EMPTY + st), ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t sK = smem_base + AttnSmem::K0 + st * AttnSmem::STAGE;
#pragma unroll
        for (int k = 0; k < DHP / 16; ++k) {
          const uint64_t da = umma_desc_k_sw128(sQ + (k >> 2) * TA_REGION) + 2u * (k & 3);
          const uint64_t db = umma_desc_k_sw128(sK + (k >> 2) * TA_REGION) + 2u * (k & 3);
          umma_f16(tmem_base + st * TA_BN, da, db, idesc_qk, k > 0 ? 1u : 0u);
        }
        umma_commit(bar(BAR_S_FULL + st));
      };
      mbar_wait(bar(BAR_Q_FULL), 0);
      if (n_tiles > 0) issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        mbar_wait(bar(BAR_P_FULL), j & 1);
        mbar_wait(bar(BAR_O_EMPTY), (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t sV = smem_base + AttnSmem::K0 + st * AttnSmem::STAGE + 2 * TA_REGION;
#pragma unroll
        for (int k = 0; k < TA_BN / 16; ++k) {
          const uint64_t da = umma_desc_k_sw128(sP + (k >> 2) * TA_REGION) + 2u * (k & 3);
          const uint64_t db = umma_desc_mn_sw128(sV, TA_REGION) + (uint64_t)((k * 2048) >> 4);
          umma_f16(tmem_base + 2 * TA_BN, da, db, idesc_pv, k > 0 ? 1u : 0u);
        }
        umma_commit(bar(BAR_O_FULL));
        umma_commit(bar(BAR_P_EMPTY));
        umma_commit(bar(BAR_KV_EMPTY + st));
        if (j + 2 < n_tiles) issue_qk(j + 2);
      }
    }
    __syncwarp();
  } else {
    // ======================= softmax: one thread per query row =======================
    const int row = warp * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float o_run[DHP];
#pragma unroll
    for (int i = 0; i < DHP; ++i) o_run[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, c_prev = 1.f;
    const int key_lim = fe.key_limit(row);    // keys with index > key_lim are masked for this row
    const float sl2 = fe.scale_log2e();
    auto fold_o = [&](int j) {                // O_run = O_run * corr_j + O_j (O_j read from TMEM)
      mbar_wait(bar(BAR_O_FULL), j & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < DHP; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row + 2 * TA_BN + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o_run[c0 + i] = fmaf(o_run[c0 + i], c_prev, __uint_as_float(v[i]));
      }
      tc_fence_before();
      mbar_arrive(bar(BAR_O_EMPTY));
    };
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      mbar_wait(bar(BAR_S_FULL + st), (j >> 1) & 1);
      tc_fence_after();
      const int k0 = j * TA_BN;
      const bool need_mask = k0 + TA_BN - 1 > key_lim;
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < TA_BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + st * TA_BN + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]);
          if (need_mask && k0 + c0 + i > key_lim) s = -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f((m_run - m_use) * sl2);
      const float msc = m_use * sl2;
      // the P buffer is free once PV_{j-1} has completed
      mbar_wait(bar(BAR_P_EMPTY), (j & 1) ^ 1);
      // pass 2: probabilities -> bf16 P tile (swizzled K-major), row sum
      float rs = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < TA_BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + st * TA_BN + c0, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
          if (need_mask) {
            if (k0 + c0 + i > key_lim) s0 = -INFINITY;
            if (k0 + c0 + i + 1 > key_lim) s1 = -INFINITY;
          }
          const float p0 = exp2f(s0 * sl2 - msc), p1 = exp2f(s1 * sl2 - msc);
          rs += p0 + p1;
          __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // four 16-B chunks of this 32-column group
          const int c = (c0 >> 3) + q;  // chunk index 0..15 along the keys
          const uint32_t addr = smem_base + AttnSmem::P + (c >> 3) * TA_REGION + sw128_off(row, c & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]),
                       "r"(pk[4 * q + 3]) : "memory");
        }
      }
      l_run = l_run * corr + rs;
      fence_proxy_async_smem();
      mbar_arrive(bar(BAR_P_FULL));
      tc_fence_before();
      mbar_arrive(bar(BAR_S_EMPTY + st));
      if (j >= 1) fold_o(j - 1);
      c_prev = corr;
      m_run = m_new;
    }
    if (n_tiles > 0) fold_o(n_tiles - 1);
    fe.store(row, o_run, m_run, l_run);
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
  __device__ VitFront(const VitAttnParams& p_) : p(p_) {
    q0 = blockIdx.x * TA_BM; h = blockIdx.y; t = blockIdx.z;
    row_stride = 3 * p.H * DH;
    const __nv_bfloat16* base = p.qkv + (long long)t * p.S * row_stride;
    gQ = base + h * DH; gK = base + (p.H + h) * DH; gV = base + (2 * p.H + h) * DH;
  }
  __device__ int n_tiles() const { return (p.S + TA_BN - 1) / TA_BN; }
  __device__ const __nv_bfloat16* any_ptr() const { return p.qkv; }
  __device__ const __nv_bfloat16* q_row(int r) const { return q0 + r < p.S ? gQ + (long long)(q0 + r) * row_stride : nullptr; }
  __device__ const __nv_bfloat16* k_row(int j, int r) const { const int k = j * TA_BN + r; return k < p.S ? gK + (long long)k * row_stride : nullptr; }
  __device__ const __nv_bfloat16* v_row(int j, int r) const { const int k = j * TA_BN + r; return k < p.S ? gV + (long long)k * row_stride : nullptr; }
  __device__ int key_limit(int) const { return p.S - 1; }
  __device__ float scale_log2e() const { return p.scale_log2e; }
  template <int DHP>
  __device__ void store(int row, const float (&o)[DHP], float, float l) const {
    const int qr = q0 + row;
    if (qr >= p.S) return;
    const float inv = 1.f / l;
    const int out_stride = (p.split_hi_lo ? 2 : 1) * p.H * DH;
    __nv_bfloat16* ob = p.out + ((long long)t * p.S + qr) * out_stride + h * DH;
#pragma unroll
    for (int c = 0; c < DH; c += 8) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = o[c + 2 * i] * inv, b = o[c + 2 * i + 1] * inv;
        __nv_bfloat162 hv = __floats2bfloat162_rn(a, b);
        hi[i] = *reinterpret_cast<uint32_t*>(&hv);
        const float2 hf = __bfloat1622float2(hv);
        __nv_bfloat162 lv = __floats2bfloat162_rn(a - hf.x, b - hf.y);
        lo[i] = *reinterpret_cast<uint32_t*>(&lv);
      }
      *reinterpret_cast<uint4*>(ob + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (p.split_hi_lo) *reinterpret_cast<uint4*>(ob + p.H * DH + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
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
  __device__ PagedFront(const PagedAttnParams& p_) : p(p_) {
    G = p.Hq / p.Hkv;
    kvh = blockIdx.x % p.Hkv;
    const int qt = blockIdx.x / p.Hkv;
    sp = blockIdx.y;
    const int st = blockIdx.z;
    q_start = p.stream_desc[st * 4 + 0]; n_q = p.stream_desc[st * 4 + 1]; kv_len = p.stream_desc[st * 4 + 2];
    table = p.block_tables + p.stream_desc[st * 4 + 3];
    R = n_q * G;
    r_base = qt * TA_BM;
    past = kv_len - n_q;
    const int last_row = min(R, r_base + TA_BM) - 1;
    const int max_pos = r_base < R ? past + last_row / G : -1;
    const int vis = r_base < R ? min((kv_len + TA_BN - 1) / TA_BN, max_pos / TA_BN + 1) : 0;
    const int per = (vis + p.n_splits - 1) / p.n_splits;
    t_begin = sp * per;
    t_end = min(vis, t_begin + per);
    if (t_end < t_begin) t_end = t_begin;
  }
  __device__ bool active() const { return r_base < R; }
  __device__ int n_tiles() const { return t_end - t_begin; }
  __device__ const __nv_bfloat16* any_ptr() const { return p.q; }
  __device__ const __nv_bfloat16* q_row(int r) const {
    const int rr = r_base + r;
    if (rr >= R) return nullptr;
    return p.q + ((long long)(q_start + rr / G) * p.Hq + kvh * G + rr % G) * DH;
  }
  __device__ const __nv_bfloat16* kv_row(int j, int r, int is_v) const {
    const int key = (t_begin + j) * TA_BN + r;
    if (key >= kv_len) return nullptr;
    const int page = table[key / PAGE];
    return p.kv_layer + ((((long long)page * 2 + is_v) * p.Hkv + kvh) * PAGE + key % PAGE) * DH;
  }
  __device__ const __nv_bfloat16* k_row(int j, int r) const { return kv_row(j, r, 0); }
  __device__ const __nv_bfloat16* v_row(int j, int r) const { return kv_row(j, r, 1); }
  // masking works on tile-local key indices j*TA_BN + c: shift the causal limit into that frame
  __device__ int key_limit(int row) const { return past + min(r_base + row, R - 1) / G - t_begin * TA_BN; }
  __device__ float scale_log2e() const { return p.scale_log2e; }
  template <int DHP>
  __device__ void store(int row, const float (&o)[DHP], float m, float l) const {
    const int rr = r_base + row;
    if (rr >= R) return;
    const long long grow = (long long)(q_start + rr / G) * p.Hq + kvh * G + rr % G;
    float* op = p.o_part + ((long long)sp * p.part_stride_rows + grow) * DH;
#pragma unroll
    for (int c = 0; c < DH; c += 4) *reinterpret_cast<float4*>(op + c) = make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]);
    float* mp = p.ml_part + ((long long)sp * p.part_stride_rows + grow) * 2;
    mp[0] = (m == -INFINITY) ? -INFINITY : m * p.scale_log2e;
    mp[1] = l;
  }
};

template <int DH, typename Params, typename Front>
__global__ void __launch_bounds__(TA_THREADS, 1) attn_tcgen05_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + AttnSmem::BARS;
  const uint32_t tmem_slot = bars + 8u * BAR_COUNT;
  const int warp = threadIdx.x >> 5;
  griddep_launch_dependents();
  if (threadIdx.x == 0) {
    constexpr uint32_t NLOAD = 32 * TA_LOADER_WARPS, NSOFT = 32 * TA_SOFTMAX_WARPS;
    mbar_init(bars + 8u * BAR_Q_FULL, NLOAD);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bars + 8u * (BAR_KV_FULL + s), NLOAD);
      mbar_init(bars + 8u * (BAR_KV_EMPTY + s), 1);
      mbar_init(bars + 8u * (BAR_S_FULL + s), 1);
      mbar_init(bars + 8u * (BAR_S_EMPTY + s), NSOFT);
    }
    mbar_init(bars + 8u * BAR_P_FULL, NSOFT);
    mbar_init(bars + 8u * BAR_P_EMPTY, 1);
    mbar_init(bars + 8u * BAR_O_FULL, 1);
    mbar_init(bars + 8u * BAR_O_EMPTY, NSOFT);
    fence_mbar_init();
  }
  // zero the operand regions once: padding chunks (head dim 72 -> 80, unused chunks) are never written by the loaders
  {
    uint4* z = reinterpret_cast<uint4*>(smem_raw + (smem_base - smem_u32(smem_raw)));
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < AttnSmem::BARS / 16; i += TA_THREADS) z[i] = zero;
  }
  if (warp == TA_SOFTMAX_WARPS + TA_LOADER_WARPS) tmem_alloc<512>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  griddep_wait();
  {
    Front fe(p);
    attention_pipeline<DH>(fe, smem_base, tmem_base);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TA_SOFTMAX_WARPS + TA_LOADER_WARPS) tmem_dealloc<512>(tmem_base);
}

}  // namespace

int launch_vit_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int T, int S, int H, int dh, int split_hi_lo, cudaStream_t s) {
  if (T <= 0) return 0;
  if (dh != 72) return -2;
  auto kern = attn_tcgen05_kernel<72, VitAttnParams, VitFront<72>>;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::TOTAL) != cudaSuccess) return -4;
    attr = true;
  }
  VitAttnParams p;
  p.qkv = qkv; p.out = out; p.S = S; p.H = H; p.split_hi_lo = split_hi_lo;
  p.scale_log2e = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  dim3 grid((S + TA_BM - 1) / TA_BM, H, T);
  if (launch_k(kern, grid, dim3(TA_THREADS), AttnSmem::TOTAL, s, p) != cudaSuccess) return -5;
  return 0;
}

int launch_kv_attention_tc_main(const __nv_bfloat16* q, const __nv_bfloat16* kv_layer, const int* stream_desc, const int* block_tables,
                                int n_streams, int max_n_q, int total_q, float* o_part, float* ml_part, int Hq, int Hkv, int n_splits,
                                cudaStream_t s) {
  auto kern = attn_tcgen05_kernel<128, PagedAttnParams, PagedFront>;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::TOTAL) != cudaSuccess) return -4;
    attr = true;
  }
  PagedAttnParams p;
  p.q = q; p.kv_layer = kv_layer; p.stream_desc = stream_desc; p.block_tables = block_tables; p.o_part = o_part; p.ml_part = ml_part;
  p.Hq = Hq; p.Hkv = Hkv; p.n_splits = n_splits; p.part_stride_rows = (long long)total_q * Hq;
  p.scale_log2e = (1.0f / sqrtf(128.f)) * 1.4426950408889634f;
  const int G = Hq / Hkv;
  dim3 grid(((max_n_q * G + TA_BM - 1) / TA_BM) * Hkv, n_splits, n_streams);
  if (launch_k(kern, grid, dim3(TA_THREADS), AttnSmem::TOTAL, s, p) != cudaSuccess) return -5;
  return 0;
}

}  // namespace mmd
