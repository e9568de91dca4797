// C-ABI of libmmduet_b200.so (declared in include/mmduet_b200.h): argument checking, workspace carving and the launch
// sequences of the SigLIP tower, the projector/pool stage and the decoder step.
#include "../../include/mmduet_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"
#include "launch.cuh"

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& s) { g_err = s; return code; }

struct Bump {  // 256-B aligned bump allocator over the caller's workspace
  uint8_t* base; int64_t cap; int64_t off = 0; bool ok = true;
  Bump(void* b, int64_t c) : base(static_cast<uint8_t*>(b)), cap(c) {}
  template <typename T> T* take(int64_t n_elems) {
    const int64_t bytes = (n_elems * (int64_t)sizeof(T) + 255) & ~int64_t(255);
    if (base != nullptr && off + bytes > cap) ok = false;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
    return p;
  }
};

inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
}  // namespace

// Launch sites of the three stage functions, by name: used for the launch counter and the optional per-stage CUDA-event
// timing (mmd_profile_start / mmd_profile_stop) that bench.py uses for the roofline of the dominant kernel.
static const char* const kTags[] = {
    "im2col", "pos_emb", "patch_embed", "ln1", "qkv", "vit_attention", "out_proj", "ln2", "fc1", "fc2",
    "gather", "proj.0+gelu", "proj.2", "tap_pool",
    "embed/concat", "input_layernorm", "qkv_proj", "qkv_finish", "kv_attention", "o_proj", "post_attention_layernorm",
    "gate_up_swiglu", "down_proj", "next_layernorm", "heads", "lm_head", "gate_up_precise", "down_precise", "final_norm_heads"};
constexpr int kNumTags = sizeof(kTags) / sizeof(kTags[0]);
static int tag_of(const char* name) {
  for (int i = 0; i < kNumTags; ++i) if (strcmp(kTags[i], name) == 0) return i;
  return -1;
}

struct ProfRec { int tag; cudaEvent_t a, b; };
struct mmd_ctx {
  int device;
  int num_sms;
  mmd::GemmContext* gemm;
  unsigned long long launches = 0;
  bool use_pdl = true;   // programmatic dependent launch across the decoder step's kernels (MMD_NO_PDL=1 disables)
  bool precise = true;   // bf16 hi+lo operands on the rows whose outputs are read (MMD_NO_PRECISE=1 disables; diagnostics)
  bool prof_on = false;
  unsigned long long prof_mask = 0;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<ProfRec> recs;
};

struct ProfScope {
  mmd_ctx* c; cudaStream_t s; cudaEvent_t b = nullptr;
  ProfScope(mmd_ctx* c_, const char* name, cudaStream_t s_, int n_kernels) : c(c_), s(s_) {
    c->launches += n_kernels;
    if (!c->prof_on) return;
    const int tag = tag_of(name);
    if (tag < 0 || !((c->prof_mask >> tag) & 1ull)) return;
    while (c->ev_pool.size() < c->ev_used + 2) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      c->ev_pool.push_back(e);
    }
    cudaEvent_t a = c->ev_pool[c->ev_used++];
    b = c->ev_pool[c->ev_used++];
    cudaEventRecord(a, s);
    c->recs.push_back({tag, a, b});
  }
  ~ProfScope() { if (b) cudaEventRecord(b, s); }
};

// Kernels launch on the CURRENT device and shared-memory attributes are kept per device: a context used while another
// device is current would run on the wrong GPU's memory, so it is refused.
static int ctx_device_mismatch(const struct mmd_ctx* c);
#define CHECK_CTX(c) do { if ((c) == nullptr) return fail(MMD_ERR_ARG, "null context"); \
                          if (int rc_ = ctx_device_mismatch(c)) return rc_; } while (0)
#define RUN(expr, what) do { int rc_ = (expr); if (rc_ != 0) return fail(rc_, std::string(what) + ": " + mmd::gemm_last_error()); } while (0)
#define RUNK(expr, what) do { int rc_ = (expr); if (rc_ != 0) return fail(rc_, std::string(what) + ": bad arguments"); } while (0)
// stage-function variants: count the launch and time it when profiling is on (needs `c` and `s` in scope)
#define PRUN(expr, what) do { ProfScope ps_(c, what, s, 1); int rc_ = (expr); if (rc_ != 0) return fail(rc_, std::string(what) + ": " + mmd::gemm_last_error()); } while (0)
#define PRUNK(expr, what) do { ProfScope ps_(c, what, s, 1); int rc_ = (expr); if (rc_ != 0) return fail(rc_, std::string(what) + ": bad arguments"); } while (0)
#define PRUNK2(expr, what) do { ProfScope ps_(c, what, s, 2); int rc_ = (expr); if (rc_ != 0) return fail(rc_, std::string(what) + ": bad arguments"); } while (0)

static int ctx_device_mismatch(const mmd_ctx* c) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess) return fail(MMD_ERR_CUDA, "cudaGetDevice failed");
  if (cur != c->device)
    return fail(MMD_ERR_ARG, "context belongs to device " + std::to_string(c->device) + " but device " + std::to_string(cur) +
                                 " is current (call cudaSetDevice / torch.cuda.set_device first)");
  return 0;
}

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MMD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}

// split-K choice for the swap-AB (weight-streaming) GEMMs: fill the SMs, few waves, not too many partial planes
static int choose_splits(int num_sms, int N, int K, int M) {
  const int bn = M <= 64 ? 64 : (M <= 128 ? 128 : 256);
  const int tiles = ((N + 127) / 128) * ((M + bn - 1) / bn);
  const int kb = (K + 63) / 64;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 8 && s <= kb; ++s) {
    const int eff = mmd::gemm_effective_splits(K, s);
    if (eff != s) continue;
    const int per = (kb + s - 1) / s;
    const long long units = (long long)tiles * s;
    const long long waves = (units + num_sms - 1) / num_sms;
    const double cost = (double)waves * (per + 6.0) + 0.5 * s;  // 6 k-blocks of fixed per-tile overhead, plane traffic
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

// `hilo`: act is a [hi | lo] pair of width 2K (row stride 2K)
static int gemm_T_partials(mmd_ctx* c, const void* act, int M, const void* w, int N, int K, int splits, float* planes,
                           cudaStream_t s, int* eff_out, bool hilo = false) {
  mmd::GemmArgs a;
  a.X = static_cast<const __nv_bfloat16*>(w); a.x_rows = N; a.ldx = K;
  a.Y = static_cast<const __nv_bfloat16*>(act); a.y_rows = M; a.ldy = hilo ? 2 * (int64_t)K : K; a.y_hilo = hilo ? 1 : 0;
  a.K = K; a.epi = mmd::EPI_T_F32; a.out = planes; a.ldo = N; a.k_splits = splits; a.split_stride = (int64_t)M * N;
  *eff_out = mmd::gemm_effective_splits(K, splits);
  return mmd::gemm_launch(c->gemm, a, s);
}

static int gemm_normal(mmd_ctx* c, const void* act, int M, const void* w, int N, int K, int64_t lda, int epi, int actfn,
                       const float* bias, void* out, int64_t ldo, cudaStream_t s) {
  mmd::GemmArgs a;
  a.X = static_cast<const __nv_bfloat16*>(act); a.x_rows = M; a.ldx = lda;
  a.Y = static_cast<const __nv_bfloat16*>(w); a.y_rows = N; a.ldy = K;
  a.K = K; a.epi = epi; a.act = actfn; a.bias = bias; a.out = out; a.ldo = ldo;
  return mmd::gemm_launch(c->gemm, a, s);
}

extern "C" {

const char* mmd_version(void) { return "mmduet_b200 0.1 (sm_100a)"; }
const char* mmd_last_error(void) { return g_err.c_str(); }

mmd_ctx* mmd_create(int device) {
  if (cudaSetDevice(device) != cudaSuccess) { g_err = "cudaSetDevice failed"; return nullptr; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { g_err = "cudaGetDeviceProperties failed"; return nullptr; }
  if (prop.major != 10) {
    g_err = "mmduet_b200 needs an sm_100a (B200) device; found sm_" + std::to_string(prop.major * 10 + prop.minor);
    return nullptr;
  }
  mmd::GemmContext* g = mmd::gemm_context_create(device);
  if (g == nullptr) { g_err = mmd::gemm_last_error(); return nullptr; }
  mmd_ctx* c = new mmd_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->gemm = g;
  const char* no_pdl = getenv("MMD_NO_PDL");
  c->use_pdl = !(no_pdl && no_pdl[0] == '1');
  const char* no_prec = getenv("MMD_NO_PRECISE");
  c->precise = !(no_prec && no_prec[0] == '1');
  return c;
}

void mmd_destroy(mmd_ctx* c) {
  if (c == nullptr) return;
  mmd::gemm_context_destroy(c->gemm);
  delete c;
}

int mmd_num_sms(mmd_ctx* c) { return c ? c->num_sms : 0; }

int mmd_set_gemm_2cta(int on) {
  mmd::g_gemm_use_2cta = on ? 1 : 0;
  return 0;
}

int mmd_set_attention_impl(int impl) {
  // bit 2 (value 4) switches the decode kernel OFF (diagnostics / A-B measurements); the low two bits select the general kernels
  if (impl < 0 || (impl & 3) > 2 || impl > 7) return fail(MMD_ERR_ARG, "mmd_set_attention_impl: 0 (mma.sync), 1 (tcgen05) or 2 (auto), +4 = no decode kernel");
  mmd::g_attention_impl = impl & 3;
  mmd::g_kv_decode = (impl & 4) ? 0 : 1;
  return 0;
}

int mmd_gemm_bf16(mmd_ctx* c, int epi, int act, const void* X, const void* X2, int64_t x_rows, int64_t ldx,
                  const void* Y, int64_t y_rows, int64_t ldy, int64_t K, const float* bias, void* out, int64_t ldo,
                  int k_splits, int64_t split_stride, void* stream) {
  CHECK_CTX(c);
  mmd::GemmArgs a;
  if (ldx < 0) { a.x_blocked = 1; ldx = K; }  // ldx = -1: X/X2 are tile-blocked weights (mmd_pack_blocked)
  a.X = static_cast<const __nv_bfloat16*>(X);
  a.X2 = static_cast<const __nv_bfloat16*>(X2);
  a.Y = static_cast<const __nv_bfloat16*>(Y);
  a.x_rows = (int)x_rows; a.y_rows = (int)y_rows; a.K = (int)K;
  a.ldx = ldx; a.ldy = ldy; a.epi = epi & 0xff; a.act = act; a.bias = bias; a.out = out; a.ldo = ldo;
  a.y_hilo = (epi & MMD_GEMM_Y_HILO) ? 1 : 0; a.out_hilo = (epi & MMD_GEMM_OUT_HILO) ? 1 : 0;
  a.k_splits = k_splits; a.split_stride = split_stride;
  RUN(mmd::gemm_launch(c->gemm, a, S(stream)), "mmd_gemm_bf16");
  return 0;
}

int mmd_gemm_splits(int64_t K, int k_splits) { return mmd::gemm_effective_splits((int)K, k_splits); }

int mmd_frame_ingest(const void* frames_bgr_hwc, int n_frames, int in_h, int in_w, void* out_rgb_chw, int res, void* stream) {
  if (frames_bgr_hwc == nullptr || out_rgb_chw == nullptr || n_frames < 0 || in_h <= 0 || in_w <= 0 || res <= 0 || res % 4 != 0)
    return fail(MMD_ERR_ARG, "mmd_frame_ingest: bad arguments (res must be a positive multiple of 4)");
  // the reference's int((short/long) * res) may truncate to 0 for extreme aspect ratios; cv2.resize raises there
  const int nw = in_w > in_h ? res : (int)(((double)in_w / (double)in_h) * res);
  const int nh = in_w > in_h ? (int)(((double)in_h / (double)in_w) * res) : res;
  if (nw < 1 || nh < 1) return fail(MMD_ERR_ARG, "mmd_frame_ingest: aspect ratio too extreme (resized side would be 0)");
  if (n_frames == 0) return 0;
  RUNK(mmd::launch_frame_ingest(static_cast<const uint8_t*>(frames_bgr_hwc), n_frames, in_h, in_w, static_cast<uint8_t*>(out_rgb_chw), res,
                                S(stream)), "mmd_frame_ingest");
  return check_launch("mmd_frame_ingest");
}

int mmd_im2col(const void* pixels, int px_dtype, int normalize, void* A, int T, int img, int patch, int k_pad, void* stream) {
  if (pixels == nullptr || A == nullptr || T < 0 || patch <= 0 || img < patch || k_pad < 3 * patch * patch)
    return fail(MMD_ERR_ARG, "mmd_im2col: bad arguments");
  if (T == 0) return 0;
  RUNK(mmd::launch_im2col(pixels, px_dtype, normalize, static_cast<__nv_bfloat16*>(A), T, 3, img, patch, k_pad, S(stream)), "mmd_im2col");
  return check_launch("mmd_im2col");
}

int mmd_layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_f32, int64_t rows, int D, float eps,
                  void* stream) {
  if (rows == 0) return 0;
  RUNK(mmd::launch_layernorm(x, gamma, beta, out, out_f32, rows, D, eps, S(stream)), "mmd_layernorm");
  return check_launch("mmd_layernorm");
}

int mmd_vit_attention(const void* qkv, void* out, int T, int S_, int H, int dh, int split_hi_lo, void* stream) {
  RUNK(mmd::launch_vit_attention(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), T, S_, H, dh, split_hi_lo, S(stream)),
       "mmd_vit_attention (head_dim must be 72)");
  return check_launch("mmd_vit_attention");
}

int mmd_resid_add_rmsnorm(float* resid, const float* partial, int n_planes, int64_t plane_stride, const float* w, void* out_bf16,
                          float* out_f32, int64_t rows, int H, float eps, void* stream) {
  RUNK(mmd::launch_resid_add_rmsnorm(resid, partial, n_planes, plane_stride, w, static_cast<__nv_bfloat16*>(out_bf16), out_f32,
                                     rows, H, eps, S(stream)), "mmd_resid_add_rmsnorm");
  return check_launch("mmd_resid_add_rmsnorm");
}

int mmd_resid_add_rmsnorm_precise(float* resid, const float* partial, int n_planes, int64_t plane_stride, const float* w, void* out_bf16,
                                  float* out_f32, int64_t rows, int H, float eps, const int* prec_of_row, const float* prec_partial,
                                  int n_prec_planes, int64_t prec_plane_stride, void* out_hilo, void* stream) {
  RUNK(mmd::launch_resid_add_rmsnorm_precise(resid, partial, n_planes, plane_stride, w, static_cast<__nv_bfloat16*>(out_bf16), out_f32, rows, H,
                                             eps, prec_of_row, prec_partial, n_prec_planes, prec_plane_stride,
                                             static_cast<__nv_bfloat16*>(out_hilo), S(stream)), "mmd_resid_add_rmsnorm_precise");
  return check_launch("mmd_resid_add_rmsnorm_precise");
}

int mmd_final_norm_heads(const float* resid, const float* partial, int n_planes, int64_t plane_stride, const float* w,
                         const int* score_rows, int n_score, const int* lm_rows, int n_lm, const float* head_w, float* logits_out,
                         float* scores_out, void* lm_x, int H, float eps, const int* prec_of_row, const float* prec_partial,
                         int n_prec_planes, int64_t prec_plane_stride, void* stream) {
  RUNK(mmd::launch_final_norm_heads(resid, partial, n_planes, plane_stride, w, score_rows, n_score, lm_rows, n_lm, head_w, logits_out,
                                    scores_out, static_cast<__nv_bfloat16*>(lm_x), H, eps, prec_of_row, prec_partial, n_prec_planes,
                                    prec_plane_stride, S(stream)), "mmd_final_norm_heads");
  return check_launch("mmd_final_norm_heads");
}

int mmd_qkv_finish(const float* partial, int n_planes, int64_t plane_stride, const float* bias, const float* rope_cos,
                   const float* rope_sin, const int* tok_pos, const int* tok_slot, void* q_out, void* kv_layer, int M, int Hq,
                   int Hkv, int dh, void* stream) {
  RUNK(mmd::launch_qkv_finish(partial, n_planes, plane_stride, bias, rope_cos, rope_sin, tok_pos, tok_slot,
                              static_cast<__nv_bfloat16*>(q_out), static_cast<__nv_bfloat16*>(kv_layer), M, Hq, Hkv, dh,
                              MMD_PAGE_TOKENS, S(stream)), "mmd_qkv_finish");
  return check_launch("mmd_qkv_finish");
}

int mmd_kv_attention_splits(mmd_ctx* c, int max_n_q, int Hq, int Hkv, int n_streams, int max_kv_len) {
  if (c == nullptr || Hkv <= 0) return 1;
  return mmd::kv_attention_pick_splits(max_n_q * (Hq / Hkv), Hkv, n_streams, max_kv_len, c->num_sms);
}

int mmd_kv_attention(mmd_ctx* c, const void* q, const void* kv_layer, const int* stream_desc, const int* block_tables,
                     int n_streams, int max_n_q, int total_q, int max_kv_len, float* o_part, float* ml_part, void* out, int Hq,
                     int Hkv, int dh, int n_splits, void* stream) {
  CHECK_CTX(c);
  if (n_splits <= 0) n_splits = mmd_kv_attention_splits(c, max_n_q, Hq, Hkv, n_streams, max_kv_len);
  RUNK(mmd::launch_kv_attention(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kv_layer), stream_desc,
                                block_tables, n_streams, max_n_q, total_q, max_kv_len, o_part, ml_part, static_cast<__nv_bfloat16*>(out),
                                Hq, Hkv, dh, MMD_PAGE_TOKENS, n_splits, S(stream)), "mmd_kv_attention (head_dim must be 128)");
  return check_launch("mmd_kv_attention");
}

int mmd_tap_pool(const void* in, int in_dtype, void* out, int out_dtype, const int* tap_idx, const float* tap_w, int T, int n_in,
                 int n_out, int max_taps, int D, int maxpool, void* stream) {
  RUNK(mmd::launch_tap_pool(in, in_dtype, out, out_dtype, tap_idx, tap_w, T, n_in, n_out, max_taps, D, maxpool, S(stream)), "mmd_tap_pool");
  return check_launch("mmd_tap_pool");
}

int mmd_heads(const float* hidden_f32, const int* rows, const float* head_w, float* logits_out, float* scores_out, int n_rows,
              int H, void* stream) {
  RUNK(mmd::launch_heads(hidden_f32, rows, head_w, logits_out, scores_out, n_rows, H, S(stream)), "mmd_heads");
  return check_launch("mmd_heads");
}

int mmd_probe_attention(const float* q, const void* kv, float* out, int T, int S_, int H, int dh, void* stream) {
  if (q == nullptr || kv == nullptr || out == nullptr) return fail(MMD_ERR_ARG, "mmd_probe_attention: null argument");
  RUNK(mmd::launch_probe_attention(q, static_cast<const __nv_bfloat16*>(kv), out, T, S_, H, dh, S(stream)), "mmd_probe_attention");
  return check_launch("mmd_probe_attention");
}

int mmd_argmax(const float* logits, int64_t V, const int64_t* penal_ids, int n_penal, float penalty, int64_t* out_id, void* stream) {
  RUNK(mmd::launch_argmax(logits, 1, 0, (int)V, reinterpret_cast<const long long*>(penal_ids), n_penal, penalty,
                          reinterpret_cast<long long*>(out_id), nullptr, S(stream)), "mmd_argmax");
  return check_launch("mmd_argmax");
}

// ---------------------------------------------------------------------------------------------------------------
// SigLIP tower
// ---------------------------------------------------------------------------------------------------------------
// Below this many token rows (one frame = 729) the residual-updating GEMMs (out_proj, fc2: N = 1152, i.e. 36 output tiles
// with a serial K loop of up to 4304) run swap-AB + split-K over all SMs instead; their fp32 partial planes, bias and
// the residual add are folded, in a fixed order, into the LayerNorm that follows.
constexpr int kVitSmallRows = 1024;
constexpr int kVitMaxSplits = 8;
struct VitBufs { __nv_bfloat16 *x, *qkv, *h; float* planes; };
static int64_t vit_carve(const mmd_vit_weights* w, int T, Bump& b, VitBufs* o) {
  const int G = w->image_size / w->patch_size;
  const int64_t M = (int64_t)T * G * G;
  o->planes = M < kVitSmallRows ? b.take<float>((int64_t)kVitMaxSplits * M * w->dim) : nullptr;
  int wide = w->mlp > w->k_pad ? w->mlp : w->k_pad;
  if (wide < 2 * w->dim) wide = 2 * w->dim;
  o->x = b.take<__nv_bfloat16>(M * w->dim);
  o->qkv = b.take<__nv_bfloat16>(M * 3 * w->dim);
  o->h = b.take<__nv_bfloat16>(M * wide);  // fc1 output; also holds the im2col matrix and the [hi|lo] attention output
  return b.off;
}

int64_t mmd_vit_workspace_bytes(const mmd_vit_weights* w, int T) {
  if (w == nullptr || T <= 0) return 0;
  Bump b(nullptr, 0);
  VitBufs o;
  return vit_carve(w, T, b, &o);
}

int mmd_vit_forward(mmd_ctx* c, const mmd_vit_weights* w, const void* pixels, int px_dtype, int normalize, int T,
                    float* resid_out, void* workspace, int64_t workspace_bytes, void* stream) {
  CHECK_CTX(c);
  if (w == nullptr || pixels == nullptr || resid_out == nullptr || workspace == nullptr || T < 0)
    return fail(MMD_ERR_ARG, "mmd_vit_forward: null argument");
  if (T == 0) return 0;
  if (w->dim % w->heads != 0 || w->dim % 8 != 0 || w->mlp % 8 != 0 || w->k_pad % 8 != 0 || w->k_pad < 3 * w->patch_size * w->patch_size)
    return fail(MMD_ERR_ARG, "mmd_vit_forward: unsupported architecture (dims must be multiples of 8)");
  Bump b(workspace, workspace_bytes);
  VitBufs buf;
  vit_carve(w, T, b, &buf);
  if (!b.ok) return fail(MMD_ERR_WORKSPACE, "mmd_vit_forward: workspace too small");
  cudaStream_t s = S(stream);
  const int G = w->image_size / w->patch_size, Sg = G * G;
  const int M = T * Sg, D = w->dim;
  const float eps = 1e-6f;
  // embeddings: resid = pos_emb (broadcast) ; resid += im2col(pixels) @ patch_w^T + patch_b
  PRUNK(mmd::launch_im2col(pixels, px_dtype, normalize, buf.h, T, 3, w->image_size, w->patch_size, w->k_pad, s), "im2col");
  PRUNK(mmd::launch_broadcast_rows(w->pos_emb, resid_out, M, Sg, D, s), "pos_emb");
  PRUN(gemm_normal(c, buf.h, M, w->patch_w, D, w->k_pad, w->k_pad, mmd::EPI_RESID_F32, 0, w->patch_b, resid_out, D, s), "patch_embed");
  const bool small = M < kVitSmallRows;
  // Live mode is a chain of ~190 kernels of 5-20 us: programmatic dependent launch overlaps each kernel's set-up with its
  // predecessor's tail.  Every kernel of the small-batch chain waits (griddepcontrol.wait) before touching global memory;
  // the CTA-pair GEMM of the batched path has no such prologue, so PDL stays off there.
  struct VitPdlGuard { VitPdlGuard(bool on) { mmd::g_use_pdl = on; } ~VitPdlGuard() { mmd::g_use_pdl = false; } } pdl_guard(c->use_pdl && small);
  // pending split-K result of the previous residual GEMM, consumed by the next LayerNorm (small-batch path only)
  const float* pend_bias = nullptr;
  int pend_planes = 0;
  auto resid_gemm = [&](const void* act, const void* wt, const float* bias, int K, const char* tag) -> int {
    if (!small) return gemm_normal(c, act, M, wt, D, K, K, mmd::EPI_RESID_F32, 0, bias, resid_out, D, s);
    int splits = choose_splits(c->num_sms, D, K, M);
    if (splits > kVitMaxSplits) splits = kVitMaxSplits;
    int eff = 1;
    const int rc = gemm_T_partials(c, act, M, wt, D, K, splits, buf.planes, s, &eff);
    pend_bias = bias;
    pend_planes = eff;
    (void)tag;
    return rc;
  };
  auto layernorm = [&](const float* g, const float* be, void* out) -> int {
    const int rc = mmd::launch_resid_add_layernorm(resid_out, buf.planes, pend_planes, (int64_t)M * D, pend_bias, g, be, out, 0, M, D, eps, s);
    pend_planes = 0;
    pend_bias = nullptr;
    return rc;
  };
  for (int l = 0; l < w->n_layers; ++l) {
    const mmd_vit_layer& L = w->layers[l];
    PRUNK(layernorm(L.ln1_w, L.ln1_b, buf.x), "ln1");
    PRUN(gemm_normal(c, buf.x, M, L.qkv_w, 3 * D, D, D, mmd::EPI_BF16, mmd::ACT_NONE, L.qkv_b, buf.qkv, 3 * D, s), "qkv");
    if (w->attn_out_split) {  // out_w is [dim, 2*dim] = [W | W]; the attention output is [hi | lo]
      PRUNK(mmd::launch_vit_attention(buf.qkv, buf.h, T, Sg, w->heads, D / w->heads, 1, s), "vit_attention");
      PRUN(resid_gemm(buf.h, L.out_w, L.out_b, 2 * D, "out_proj"), "out_proj");
    } else {
      PRUNK(mmd::launch_vit_attention(buf.qkv, buf.x, T, Sg, w->heads, D / w->heads, 0, s), "vit_attention");
      PRUN(resid_gemm(buf.x, L.out_w, L.out_b, D, "out_proj"), "out_proj");
    }
    PRUNK(layernorm(L.ln2_w, L.ln2_b, buf.x), "ln2");
    PRUN(gemm_normal(c, buf.x, M, L.fc1_w, w->mlp, D, D, mmd::EPI_BF16, mmd::ACT_GELU_TANH, L.fc1_b, buf.h, w->mlp, s), "fc1");
    PRUN(resid_gemm(buf.h, L.fc2_w, L.fc2_b, w->mlp, "fc2"), "fc2");
  }
  if (pend_planes > 0) PRUNK(layernorm(nullptr, nullptr, nullptr), "ln1");   // the last fc2: residual update only
  return check_launch("mmd_vit_forward");
}

// ---------------------------------------------------------------------------------------------------------------
// projector + pooling
// ---------------------------------------------------------------------------------------------------------------
struct ProjBufs { __nv_bfloat16 *g, *a; float* b; };
static bool proj_pooled(const mmd_projector_weights* w) {
  return (w->pool_group == 4 || w->pool_group == 16) && w->hilo && !w->maxpool && w->pool_gather_idx != nullptr && w->pool_row_w != nullptr;
}
static int64_t proj_carve(const mmd_projector_weights* w, int T, Bump& b, ProjBufs* o) {
  const int f = w->hilo ? 2 : 1;
  if (proj_pooled(w)) {
    o->g = b.take<__nv_bfloat16>((int64_t)T * w->n_out * w->pool_group * w->vit_dim * 2);   // tap-major gathered tokens [hi | lo]
    o->a = b.take<__nv_bfloat16>((int64_t)T * w->n_out * w->hidden * 2);                    // pooled GELU activations [hi | lo]
    o->b = nullptr;
    return b.off;
  }
  o->g = b.take<__nv_bfloat16>((int64_t)T * w->n_gather * w->vit_dim * f);
  o->a = b.take<__nv_bfloat16>((int64_t)T * w->n_gather * w->hidden * f);
  o->b = b.take<float>((int64_t)T * w->n_gather * w->hidden);
  return b.off;
}
int64_t mmd_projector_workspace_bytes(const mmd_projector_weights* w, int T) {
  if (w == nullptr || T <= 0) return 0;
  Bump b(nullptr, 0);
  ProjBufs o;
  return proj_carve(w, T, b, &o);
}

int mmd_projector_pool(mmd_ctx* c, const mmd_projector_weights* w, const float* vit_resid, int T, void* out, int out_dtype,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  CHECK_CTX(c);
  if (w == nullptr || vit_resid == nullptr || out == nullptr || workspace == nullptr || T < 0)
    return fail(MMD_ERR_ARG, "mmd_projector_pool: null argument");
  if (out_dtype != MMD_DT_BF16 && out_dtype != MMD_DT_F32) return fail(MMD_ERR_ARG, "mmd_projector_pool: out dtype must be bf16 or f32");
  if (T == 0) return 0;
  if (w->vit_dim % 8 != 0 || w->hidden % 8 != 0) return fail(MMD_ERR_ARG, "mmd_projector_pool: dims must be multiples of 8");
  Bump b(workspace, workspace_bytes);
  ProjBufs buf;
  proj_carve(w, T, b, &buf);
  if (!b.ok) return fail(MMD_ERR_WORKSPACE, "mmd_projector_pool: workspace too small");
  cudaStream_t s = S(stream);
  if (proj_pooled(w)) {
    // gather tap-major -> Linear1 + GELU + pooling taps in the epilogue -> Linear2 on the pooled rows, straight into `out`
    const int G = w->pool_group, Mg = T * w->n_out * G, Mo = T * w->n_out, H = w->hidden, D = w->vit_dim;
    PRUNK(mmd::launch_gather_rows_f32_to_bf16(vit_resid, w->pool_gather_idx, buf.g, T, w->n_src_tokens, w->n_out * G, D, 1, s), "gather");
    {
      mmd::GemmArgs a;
      a.X = buf.g; a.x_rows = Mg; a.ldx = 2 * (int64_t)D; a.Y = static_cast<const __nv_bfloat16*>(w->w1); a.y_rows = H; a.ldy = 2 * (int64_t)D;
      a.K = 2 * D; a.epi = mmd::EPI_BF16_HILO_POOL; a.act = mmd::ACT_GELU_ERF; a.bias = w->b1; a.out = buf.a; a.ldo = 2 * (int64_t)H;
      a.row_w = w->pool_row_w; a.row_w_period = w->n_out * G; a.pool_group = G;
      PRUN(mmd::gemm_launch(c->gemm, a, s), "proj.0+gelu");
    }
    {
      mmd::GemmArgs a;
      a.X = buf.a; a.x_rows = Mo; a.ldx = 2 * (int64_t)H; a.Y = static_cast<const __nv_bfloat16*>(w->w2); a.y_rows = H; a.ldy = 2 * (int64_t)H;
      a.K = 2 * H; a.epi = out_dtype == MMD_DT_F32 ? mmd::EPI_F32 : mmd::EPI_BF16; a.act = mmd::ACT_NONE; a.bias = w->b2; a.out = out; a.ldo = H;
      a.force_1cta = 1;          // per-thread stores: `out` may be another GPU's memory (PeerStoreEncoder)
      PRUN(mmd::gemm_launch(c->gemm, a, s), "proj.2");
    }
    return check_launch("mmd_projector_pool");
  }
  const int Mg = T * w->n_gather, H = w->hidden, f = w->hilo ? 2 : 1;
  // Only the source tokens the pooling reads go through the projector (169 of 729 for bilinear 27->7).  With `hilo` the
  // GEMM operands are bf16 hi+lo pairs (K doubled against [W | W]), Linear2 keeps its fp32 accumulator and the pooling
  // combines in fp32: the frame tokens are rounded to bf16 exactly once, at the output.
  PRUNK(mmd::launch_gather_rows_f32_to_bf16(vit_resid, w->gather_idx, buf.g, T, w->n_src_tokens, w->n_gather, w->vit_dim, w->hilo, s), "gather");
  PRUN(gemm_normal(c, buf.g, Mg, w->w1, H, f * w->vit_dim, f * w->vit_dim, w->hilo ? mmd::EPI_BF16_HILO : mmd::EPI_BF16, mmd::ACT_GELU_ERF,
                  w->b1, buf.a, f * H, s), "proj.0+gelu");
  PRUN(gemm_normal(c, buf.a, Mg, w->w2, H, f * H, f * H, mmd::EPI_F32, mmd::ACT_NONE, w->b2, buf.b, H, s), "proj.2");
  PRUNK(mmd::launch_tap_pool(buf.b, mmd::DT_F32, out, out_dtype == MMD_DT_F32 ? mmd::DT_F32 : mmd::DT_BF16, w->tap_idx, w->tap_w, T,
                            w->n_gather, w->n_out, w->max_taps, H, w->maxpool, s), "tap_pool");
  return check_launch("mmd_projector_pool");
}

// ---------------------------------------------------------------------------------------------------------------
// decoder step
// ---------------------------------------------------------------------------------------------------------------
struct DecBufs {
  float *resid, *planes, *graw, *o_part, *ml_part;
  __nv_bfloat16 *x, *q, *attn, *h, *lm_x, *xp2, *hp2;
};
constexpr int kMaxPrecRows = 128;   // rows carried as bf16 hi+lo pairs (one <= 128-token tile of the swap-AB GEMM)
constexpr int kMaxSplits = 8;
constexpr int kMaxAttnSplits = 32;
constexpr int kMaxDecodeSplits = 64;   // decode kernel (<= 2 tokens per stream): one stream over all SMs
static int64_t dec_carve(const mmd_dec_weights* w, int max_tokens, int max_lm_rows, Bump& b, DecBufs* o) {
  const int64_t M = max_tokens;
  const int H = w->hidden, QD = w->q_heads * w->head_dim, NQKV = (w->q_heads + 2 * w->kv_heads) * w->head_dim;
  const int nmax = NQKV > H ? NQKV : H;
  // passes above 128 tokens append 2 rows (hi, lo) per precise row to the activation matrices of q/k/v, gate/up and down
  const int64_t Mx = M > kMaxPrecRows ? M + 2 * kMaxPrecRows : M;
  o->resid = b.take<float>(M * H);
  o->planes = b.take<float>((int64_t)kMaxSplits * Mx * nmax);
  o->graw = b.take<float>(M > kMaxPrecRows ? (int64_t)2 * kMaxPrecRows * 2 * w->mlp : 1);   // raw (gate, up) of the appended rows
  o->xp2 = b.take<__nv_bfloat16>((int64_t)kMaxPrecRows * 2 * H);
  o->hp2 = b.take<__nv_bfloat16>((int64_t)kMaxPrecRows * 2 * w->mlp);
  o->x = b.take<__nv_bfloat16>(Mx * H);
  o->q = b.take<__nv_bfloat16>(M * QD);
  o->attn = b.take<__nv_bfloat16>(M * QD);
  o->h = b.take<__nv_bfloat16>(Mx * w->mlp);
  const int64_t part_rows = kMaxAttnSplits * M > 2 * kMaxDecodeSplits ? kMaxAttnSplits * M : 2 * kMaxDecodeSplits;
  o->o_part = b.take<float>(part_rows * QD);
  o->ml_part = b.take<float>(part_rows * w->q_heads * 2);
  o->lm_x = b.take<__nv_bfloat16>((int64_t)(max_lm_rows > 0 ? max_lm_rows : 1) * H);
  return b.off;
}

int64_t mmd_decoder_workspace_bytes(mmd_ctx*, const mmd_dec_weights* w, int max_tokens, int max_lm_rows) {
  if (w == nullptr || max_tokens <= 0) return 0;
  Bump b(nullptr, 0);
  DecBufs o;
  return dec_carve(w, max_tokens, max_lm_rows, b, &o);
}

int mmd_decoder_step(mmd_ctx* c, const mmd_dec_weights* w, const mmd_kv_pool* pool, const mmd_step* st, void* workspace,
                     int64_t workspace_bytes, void* stream) {
  CHECK_CTX(c);
  if (w == nullptr || pool == nullptr || st == nullptr || workspace == nullptr) return fail(MMD_ERR_ARG, "mmd_decoder_step: null argument");
  const int M = st->n_tokens;
  if (M <= 0) return 0;
  if (w->head_dim != 128 || w->q_heads % w->kv_heads != 0 || w->hidden % 8 != 0 || w->mlp % 8 != 0)
    return fail(MMD_ERR_ARG, "mmd_decoder_step: unsupported architecture (head_dim 128, dims multiples of 8)");
  if (st->n_streams <= 0 || st->stream_desc == nullptr || st->block_tables == nullptr || st->tok_pos == nullptr ||
      st->tok_slot == nullptr || (st->src_row == nullptr && st->resid_in == nullptr))
    return fail(MMD_ERR_ARG, "mmd_decoder_step: incomplete step description");
  if (st->resid_in == nullptr && w->embed == nullptr) return fail(MMD_ERR_ARG, "mmd_decoder_step: no embedding table and no resid_in");
  if (st->resid_out == nullptr && (st->n_score_rows > 0 || st->n_lm_rows > 0) && w->final_norm_w == nullptr)
    return fail(MMD_ERR_ARG, "mmd_decoder_step: this stage has no final norm (pass resid_out)");
  if (st->max_kv_len > w->max_pos) return fail(MMD_ERR_ARG, "mmd_decoder_step: context exceeds the RoPE table (max_pos)");
  if (st->n_lm_rows > 0 && (w->lm_head == nullptr || st->lm_logits_out == nullptr || st->lm_rows == nullptr))
    return fail(MMD_ERR_ARG, "mmd_decoder_step: lm rows requested without lm_head / output buffer");
  Bump b(workspace, workspace_bytes);
  DecBufs buf;
  dec_carve(w, M, st->n_lm_rows, b, &buf);
  if (!b.ok) return fail(MMD_ERR_WORKSPACE, "mmd_decoder_step: workspace too small");
  cudaStream_t s = S(stream);
  const int H = w->hidden, Hq = w->q_heads, Hkv = w->kv_heads, dh = w->head_dim;
  const int QD = Hq * dh, NQKV = (Hq + 2 * Hkv) * dh, I = w->mlp;

  struct PdlGuard { PdlGuard(bool on) { mmd::g_use_pdl = on; } ~PdlGuard() { mmd::g_use_pdl = false; } } pdl_guard(c->use_pdl);
  // "Precise rows": the scores are read at a handful of rows (one per frame) and their error is dominated by the bf16 rounding
  // of the MLP operands ON THOSE ROWS (tools/noise_floor_decoder.py: x_mlp 0.012, h 0.007 of a 0.018 total; other rows' noise
  // averages out through attention).  Those rows therefore carry their GEMM operands as bf16 hi+lo pairs:
  //   all_prec (M <= 128, single-frame / query / generation steps): every row, inside the main swap-AB kernels (the step is
  //            HBM-bound, the second MMA per weight tile is free); q/k/v, gate/up and down operands are [hi | lo], 2K wide.
  //   app_prec (M > 128): the <= 128 listed rows are APPENDED to the activation matrices of q/k/v, gate/up and down as two
  //            extra rows each (row M + j = hi, row M + P + j = lo), so they ride through the main GEMMs (weights streamed once,
  //            usually inside the last, partly filled token tile); a linear layer's precise output is the sum of the two rows'
  //            outputs, taken by the consumer (RoPE / norm kernels); gate/up writes the appended rows' raw pre-activations and a
  //            small kernel applies SwiGLU to gate_hi + gate_lo, up_hi + up_lo.
  const bool prec_ok = c->precise && H % 64 == 0 && I % 64 == 0;
  const bool all_prec = prec_ok && M <= kMaxPrecRows;
  const int P = st->n_prec_rows;
  const bool app_prec = prec_ok && !all_prec && P > 0 && P <= kMaxPrecRows && st->prec_rows != nullptr && st->prec_of_row != nullptr;
  const int Mt = app_prec ? M + 2 * P : M;                   // activation rows of the GEMMs that carry the appended rows
  __nv_bfloat16* xmain = all_prec ? nullptr : buf.x;         // plain bf16 normalised rows (unused when every row is [hi | lo])
  const int* prec_of_row = app_prec ? st->prec_of_row : nullptr;
  // inputs_embeds = cat(embed_tokens(prefix ids), frame tokens) -> fp32 residual stream   (test/inference.py:235-238)
  if (st->resid_in != nullptr) {   // a later stage of a layer pipeline: the residual stream arrives from the previous stage
    if (cudaMemcpyAsync(buf.resid, st->resid_in, (size_t)M * H * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return fail(MMD_ERR_CUDA, "mmd_decoder_step: copy of resid_in failed");
  } else {
    PRUNK(mmd::launch_gather_rows_bf16_to_f32(static_cast<const __nv_bfloat16*>(w->embed), static_cast<const __nv_bfloat16*>(st->frame_tokens),
                                             st->src_row, buf.resid, M, H, s), "embed/concat");
  }
  PRUNK(mmd::launch_resid_add_rmsnorm_precise(buf.resid, nullptr, 0, 0, w->layers[0].ln1_w, xmain, nullptr, M, H, w->rms_eps, prec_of_row,
                                             nullptr, 0, 0, all_prec ? buf.xp2 : nullptr, s, 0, app_prec ? 1 : 0, P), "input_layernorm");
  const int s_qkv = choose_splits(c->num_sms, NQKV, H, Mt);
  const int s_o = choose_splits(c->num_sms, H, QD, M);
  const int s_down = choose_splits(c->num_sms, H, I, Mt);
  int attn_splits = mmd::kv_attention_pick_splits(st->max_n_q * (Hq / Hkv), Hkv, st->n_streams, st->max_kv_len, c->num_sms);
  {   // the partial buffers hold kMaxAttnSplits x M rows, or kMaxDecodeSplits splits of <= 16 rows (dec_carve)
    const int64_t cap_rows = kMaxAttnSplits * (int64_t)M > 2 * kMaxDecodeSplits ? kMaxAttnSplits * (int64_t)M : 2 * kMaxDecodeSplits;
    while (attn_splits > 1 && (int64_t)attn_splits * M > cap_rows) --attn_splits;
    if (attn_splits > kMaxDecodeSplits) attn_splits = kMaxDecodeSplits;
  }
  // Swap-AB + split-K (weights on the 128 UMMA rows, all SMs busy) wins up to ~1200 tokens per pass on this model
  // (measured: 8/10/24-frame passes, profiles/r01_decoder_paths.md); the normal-orientation CTA-pair path (UMMA M = 256,
  // TMA-store epilogue, interleaved gate/up + pairwise SwiGLU) only pays off for much larger multi-stream batches.
  const bool big = M >= 2048;      // (kernel choices follow the pass size M, not M + appended rows: measured thresholds)
  // q/k/v, o and down (plain fp32 outputs) already win in the normal-orientation CTA-pair kernel from 1024 tokens
  // (measured at 1960 tokens: -5 %, -5 %, -10 %: no split-K planes, whole waves of 256x256 tiles); gate/up keeps the
  // swap-AB SwiGLU epilogue up to 2048.
  const bool pair_qkv = M >= 1024, pair_o = M >= 1024, pair_down = M >= 1024;
  auto gemm_big_f32 = [&](const void* act, int rows, const void* wt, int N, int K, int splits, int* eff) -> int {
    mmd::GemmArgs a;
    a.X = static_cast<const __nv_bfloat16*>(act); a.x_rows = rows; a.ldx = K;
    a.Y = static_cast<const __nv_bfloat16*>(wt); a.y_rows = N; a.ldy = K; a.K = K;
    a.epi = mmd::EPI_F32; a.out = buf.planes; a.ldo = N; a.k_splits = splits; a.split_stride = (int64_t)rows * N; a.force_2cta = 1;
    *eff = mmd::gemm_effective_splits(K, splits);
    return mmd::gemm_launch(c->gemm, a, s);
  };
  // swap-AB SwiGLU on hi+lo activations [rows, 2H] -> hi+lo outputs [rows, 2I] (gate rows / up rows of the interleaved matrix)
  auto gate_up_hilo = [&](const __nv_bfloat16* gu, const __nv_bfloat16* act2, int rows, __nv_bfloat16* out2) -> int {
    mmd::GemmArgs a;
    a.X = gu; a.X2 = gu + H; a.x_rows = I; a.ldx = 2 * (int64_t)H; a.Y = act2; a.y_rows = rows; a.ldy = 2 * (int64_t)H; a.K = H;
    a.epi = mmd::EPI_T_SWIGLU; a.out = out2; a.ldo = 2 * (int64_t)I; a.y_hilo = 1; a.out_hilo = 1;
    return mmd::gemm_launch(c->gemm, a, s);
  };
  int eff = 1;   // split-K planes of the pending residual update
  for (int l = 0; l < w->n_layers; ++l) {
    const mmd_dec_layer& L = w->layers[l];
    if (all_prec) PRUN(gemm_T_partials(c, buf.xp2, M, L.qkv_w, NQKV, H, s_qkv, buf.planes, s, &eff, true), "qkv_proj");
    else if (pair_qkv) PRUN(gemm_big_f32(buf.x, Mt, L.qkv_w, NQKV, H, 1, &eff), "qkv_proj");
    else PRUN(gemm_T_partials(c, buf.x, Mt, L.qkv_w, NQKV, H, s_qkv, buf.planes, s, &eff), "qkv_proj");
    __nv_bfloat16* kv_layer = static_cast<__nv_bfloat16*>(pool->pool) + (int64_t)l * pool->layer_stride;
    PRUNK(mmd::launch_qkv_finish(buf.planes, eff, (int64_t)Mt * NQKV, L.qkv_b, w->rope_cos, w->rope_sin, st->tok_pos, st->tok_slot,
                                buf.q, kv_layer, M, Hq, Hkv, dh, MMD_PAGE_TOKENS, s, prec_of_row, P), "qkv_finish");
    PRUNK2(mmd::launch_kv_attention(buf.q, kv_layer, st->stream_desc, st->block_tables, st->n_streams, st->max_n_q, M, st->max_kv_len, buf.o_part,
                                  buf.ml_part, buf.attn, Hq, Hkv, dh, MMD_PAGE_TOKENS, attn_splits, s), "kv_attention");
    if (pair_o) PRUN(gemm_big_f32(buf.attn, M, L.o_w, H, QD, 1, &eff), "o_proj");
    else PRUN(gemm_T_partials(c, buf.attn, M, L.o_w, H, QD, s_o, buf.planes, s, &eff), "o_proj");
    // resid += o_proj; x = RMSNorm(resid): bf16 rows (+ the appended hi / lo rows of the precise rows), or [hi | lo] for every row
    PRUNK(mmd::launch_resid_add_rmsnorm_precise(buf.resid, buf.planes, eff, (int64_t)M * H, L.ln2_w, xmain, nullptr, M, H, w->rms_eps,
                                               prec_of_row, nullptr, 0, 0, all_prec ? buf.xp2 : nullptr, s, 0, app_prec ? 1 : 0, P),
          "post_attention_layernorm");
    const __nv_bfloat16* gu = static_cast<const __nv_bfloat16*>(L.gate_up_w);
    if (all_prec) {
      PRUN(gate_up_hilo(gu, buf.xp2, M, buf.hp2), "gate_up_swiglu");
      PRUN(gemm_T_partials(c, buf.hp2, M, L.down_w, H, I, s_down, buf.planes, s, &eff, true), "down_proj");
    } else {
      mmd::GemmArgs a;
      if (big) {   // interleaved (gate, up) columns, SwiGLU on adjacent accumulator columns
        a.X = buf.x; a.x_rows = Mt; a.ldx = H; a.Y = gu; a.y_rows = 2 * I; a.ldy = H; a.K = H;
        a.epi = mmd::EPI_SWIGLU_PAIR; a.out = buf.h; a.ldo = I; a.force_2cta = 1;
      } else if (M > 1024 && I % 2 == 0) {   // (at 490 tokens the two-accumulator form is still 7 % faster: 4 re-reads only)
        // swap-AB on the interleaved matrix as ONE operand: a single accumulator per 128 weight rows (64 gate/up pairs on
        // adjacent TMEM lanes) leaves room for 256-token tiles, i.e. UMMA N = 256 instead of 2 x 128 (ncu at 1960 tokens:
        // tensor pipe 87.8 % active vs 80.6 % with two accumulators; L2->SM traffic is the same 6.5 GB either way)
        a.X = gu; a.x_rows = 2 * I; a.ldx = H; a.Y = buf.x; a.y_rows = Mt; a.ldy = H; a.K = H;
        a.epi = mmd::EPI_T_SWIGLU_IL; a.out = buf.h; a.ldo = I;
      } else {     // swap-AB: gate rows and up rows of the interleaved matrix as two strided operands
        a.X = gu; a.X2 = gu + H; a.x_rows = I; a.ldx = 2 * (int64_t)H; a.Y = buf.x; a.y_rows = Mt; a.ldy = H; a.K = H;
        a.epi = mmd::EPI_T_SWIGLU; a.out = buf.h; a.ldo = I;
      }
      if (app_prec) { a.raw_out = buf.graw; a.raw_from = M; a.ld_raw = 2 * (int64_t)I; }
      PRUN(mmd::gemm_launch(c->gemm, a, s), "gate_up_swiglu");
      if (app_prec)   // h rows M .. M + 2P: SwiGLU of (gate_hi + gate_lo, up_hi + up_lo) as a hi row and a lo row
        PRUNK(mmd::launch_swiglu_from_raw(buf.graw, 2 * (int64_t)I, buf.h + (int64_t)M * I, I, P, I, s), "gate_up_precise");
      if (pair_down) PRUN(gemm_big_f32(buf.h, Mt, L.down_w, H, I, 3, &eff), "down_proj");
      else PRUN(gemm_T_partials(c, buf.h, Mt, L.down_w, H, I, s_down, buf.planes, s, &eff), "down_proj");
    }
    if (l + 1 < w->n_layers)   // resid += down_proj (precise rows: their hi + lo copies' outputs); x = RMSNorm(resid) for the next layer
      PRUNK(mmd::launch_resid_add_rmsnorm_precise(buf.resid, buf.planes, eff, (int64_t)Mt * H, w->layers[l + 1].ln1_w, xmain, nullptr, M, H,
                                                 w->rms_eps, prec_of_row, app_prec ? buf.planes + (int64_t)M * H : nullptr, eff,
                                                 (int64_t)Mt * H, all_prec ? buf.xp2 : nullptr, s, (int64_t)P * H, app_prec ? 1 : 0, P),
            "next_layernorm");
  }
  if (st->resid_out != nullptr) {
    // not the last stage: fold the pending down_proj planes into the residual stream (precise rows: their hi + lo copies'
    // outputs) and hand it on; model.norm, the heads and lm_head belong to the last stage
    PRUNK(mmd::launch_resid_add_rmsnorm_precise(buf.resid, buf.planes, eff, (int64_t)Mt * H, nullptr, nullptr, nullptr, M, H, w->rms_eps,
                                               prec_of_row, app_prec ? buf.planes + (int64_t)M * H : nullptr, eff, (int64_t)Mt * H, nullptr,
                                               s, (int64_t)P * H, 0, P), "next_layernorm");
    if (cudaMemcpyAsync(st->resid_out, buf.resid, (size_t)M * H * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return fail(MMD_ERR_CUDA, "mmd_decoder_step: copy to resid_out failed");
    return check_launch("mmd_decoder_step");
  }
  // final norm on the rows that are read only, with the informative/relevance heads as its epilogue (+ bf16 rows for lm_head)
  if (st->n_score_rows > 0 && (st->score_rows == nullptr || st->head_logits_out == nullptr || st->scores_out == nullptr || w->heads_w == nullptr))
    return fail(MMD_ERR_ARG, "mmd_decoder_step: score rows requested without buffers");
  if (st->n_score_rows > 0 || st->n_lm_rows > 0)
    PRUNK(mmd::launch_final_norm_heads(buf.resid, buf.planes, eff, (int64_t)Mt * H, w->final_norm_w, st->score_rows, st->n_score_rows,
                                      st->lm_rows, st->n_lm_rows, w->heads_w, st->head_logits_out, st->scores_out, buf.lm_x, H, w->rms_eps,
                                      prec_of_row, app_prec ? buf.planes + (int64_t)M * H : nullptr, eff, (int64_t)Mt * H, s, (int64_t)P * H),
          "final_norm_heads");
  if (st->n_lm_rows > 0) {
    int e2 = 1;
    PRUN(gemm_T_partials(c, buf.lm_x, st->n_lm_rows, w->lm_head, w->vocab, H, 1, st->lm_logits_out, s, &e2), "lm_head");
  }
  return check_launch("mmd_decoder_step");
}

int mmd_profile_num_tags(void) { return kNumTags; }
const char* mmd_profile_tag_name(int i) { return (i >= 0 && i < kNumTags) ? kTags[i] : ""; }
unsigned long long mmd_launch_count(mmd_ctx* c) { return c ? c->launches : 0; }

int mmd_profile_start(mmd_ctx* c, const char* tags_csv) {
  CHECK_CTX(c);
  c->prof_mask = 0;
  std::string csv = tags_csv ? tags_csv : "all";
  if (csv == "all") c->prof_mask = ~0ull;
  else {
    size_t pos = 0;
    while (pos <= csv.size()) {
      size_t e = csv.find(',', pos);
      if (e == std::string::npos) e = csv.size();
      const int t = tag_of(csv.substr(pos, e - pos).c_str());
      if (t < 0) return fail(MMD_ERR_ARG, "mmd_profile_start: unknown tag " + csv.substr(pos, e - pos));
      c->prof_mask |= 1ull << t;
      pos = e + 1;
    }
  }
  c->recs.clear();
  c->ev_used = 0;
  c->prof_on = true;
  return 0;
}

int mmd_profile_stop(mmd_ctx* c, float* ms_sum, int* counts) {
  CHECK_CTX(c);
  c->prof_on = false;
  for (int i = 0; i < kNumTags; ++i) { if (ms_sum) ms_sum[i] = 0.f; if (counts) counts[i] = 0; }
  for (const ProfRec& r : c->recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) return fail(MMD_ERR_CUDA, "mmd_profile_stop: event sync failed");
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) return fail(MMD_ERR_CUDA, "mmd_profile_stop: elapsed failed");
    if (ms_sum) ms_sum[r.tag] += ms;
    if (counts) counts[r.tag] += 1;
  }
  c->recs.clear();
  c->ev_used = 0;
  return 0;
}

}  // extern "C"
