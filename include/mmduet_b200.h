/* mmduet_b200 — C ABI of the B200-native per-frame streaming hot path of MMDuet.
 *
 * The reference (yellow-binary-tree/MMDuet) is pure Python and has no FFI; its boundary is the Python surface in
 * models/vision_live.py, models/modeling_live.py, models/live_llava/video_head_live_llava_qwen.py and
 * test/inference.py.  The functions below are what a ctypes binding underneath that surface calls (see
 * INTEGRATION.md); each one names the reference call site whose arithmetic it replaces.
 *
 * Conventions: every pointer is a caller-owned CUDA DEVICE pointer unless marked "host"; sizes/strides are in
 * ELEMENTS; `stream` is a cudaStream_t passed as void*; calls are asynchronous with respect to the host, allocate
 * nothing on the device (the caller provides workspaces sized by the *_workspace_bytes functions) and never
 * synchronise; the return value is 0 or a negative MMD_ERR_* code with a thread-local message available from
 * mmd_last_error().  There is no CPU fallback: mmd_create fails on anything that is not an sm_100 device.
 * Kernels launch on the CURRENT device: every call that takes an mmd_ctx refuses (MMD_ERR_ARG) to run while a device other
 * than the context's is current, instead of launching on the wrong GPU.
 */
#ifndef MMDUET_B200_H_
#define MMDUET_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MMD_API __attribute__((visibility("default")))
#else
#define MMD_API
#endif

#define MMD_OK 0
#define MMD_ERR_ARG (-2)
#define MMD_ERR_WORKSPACE (-3)
#define MMD_ERR_CUDA (-5)

#define MMD_DT_U8 0
#define MMD_DT_BF16 1
#define MMD_DT_F32 2

#define MMD_PAGE_TOKENS 64 /* tokens per KV page */

typedef struct mmd_ctx mmd_ctx;

MMD_API const char* mmd_version(void);
MMD_API const char* mmd_last_error(void);
MMD_API mmd_ctx* mmd_create(int device);
MMD_API void mmd_destroy(mmd_ctx*);
MMD_API int mmd_num_sms(mmd_ctx*);
/* Attention kernels: 0 = mma.sync flash attention (csrc/vit_attention.cu, csrc/kv_attention.cu); 1 = tcgen05.mma / TMEM
 * flash attention (csrc/attn_tcgen05.cu); 2 (default) = auto: tcgen05 for the ViT and for decoder steps with >= 1024 stacked
 * query rows or >= 2k context, mma.sync for short single-frame steps.  Both are parity-tested.  Process-wide. */
MMD_API int mmd_set_attention_impl(int impl);
/* 1 (default): large-M GEMMs (ViT, projector) run on CTA pairs (tcgen05.mma.cta_group::2, UMMA M = 256); 0: single-CTA
 * kernel everywhere.  Process-wide. */
MMD_API int mmd_set_gemm_2cta(int on);

/* ---------------------------------------------------------------------------------------------------------------
 * GEMM (tcgen05.mma, TMA operands, TMEM accumulators) — every nn.Linear on the path.
 * D[i, j] = sum_k X[i, k] * Y[j, k], bf16 operands (both K-major), fp32 accumulation.
 * Normal epilogues: X = activations [M,K], Y = weights [N,K], out[M,N].  T epilogues (swap-AB for small M): X (and
 * X2) = weights [N,K], Y = activations [M,K], out[M,N]; MMD_EPI_T_F32 writes split-K fp32 planes `split_stride`
 * elements apart (mmd_gemm_splits gives the effective plane count for a requested split).
 * Replaces cuBLAS under: SigLIP q/k/v/out/fc1/fc2 + patch-embed Conv2d (video_head_live_llava_qwen.py:96-98),
 * mm_projector (:90-91), Qwen2 q/k/v/o/gate/up/down + lm_head (:141-155).
 * ------------------------------------------------------------------------------------------------------------- */
#define MMD_EPI_BF16 0      /* out_bf16 = act(acc + bias[n])                       */
#define MMD_EPI_RESID_F32 1 /* out_f32 += acc + bias[n]      (fp32 residual stream) */
#define MMD_EPI_T_F32 2     /* out_f32[split][m][n] = acc                          */
#define MMD_EPI_T_SWIGLU 3  /* out_bf16[m][n] = silu(acc_gate) * acc_up            */
#define MMD_EPI_F32 4       /* out_f32 = acc + bias[n]                             */
#define MMD_EPI_SWIGLU_PAIR 6 /* interleaved gate/up weight rows: out_bf16[m,j] = silu(acc[2j]) * acc[2j+1] (CTA-pair kernel) */
#define MMD_EPI_T_SWIGLU_IL 7 /* swap-AB, X rows interleaved (2j = gate_j, 2j+1 = up_j): out_bf16[m][j] = silu(acc[2j]) * acc[2j+1];
                                out is [M, x_rows / 2]; one accumulator, 256-token tiles (decoder gate/up above 128 tokens) */
#define MMD_EPI_BF16_HILO 5 /* v = act(acc + bias[n]); out[m,n] = bf16(v), out[m,N+n] = bf16(v - bf16(v)) */
#define MMD_EPI_BF16_HILO_POOL 8 /* pooling epilogue (see mmd_projector_pool); not reachable through mmd_gemm_bf16 */
/* OR-ed into `epi` (swap-AB epilogues MMD_EPI_T_F32 / MMD_EPI_T_SWIGLU, y_rows <= 128, K % 64 == 0):
 * Y_HILO: Y is a bf16 hi+lo pair [hi | lo] of width 2K (ldy >= 2K); both halves are multiplied with every weight tile into
 * the same fp32 accumulator, so the weights are streamed once and the activation rounding error drops from 2^-9 to 2^-17.
 * OUT_HILO (MMD_EPI_T_SWIGLU): the output row is [bf16(v) | bf16(v - bf16(v))], lo at column x_rows (ldo >= 2 * x_rows).
 * Used for the decoder rows whose scores are read ("precise rows"). */
#define MMD_GEMM_Y_HILO 0x100
#define MMD_GEMM_OUT_HILO 0x200
#define MMD_ACT_NONE 0
#define MMD_ACT_GELU_TANH 1
#define MMD_ACT_GELU_ERF 2
MMD_API int mmd_gemm_bf16(mmd_ctx*, int epi, int act, const void* X, const void* X2, int64_t x_rows, int64_t ldx,
                          const void* Y, int64_t y_rows, int64_t ldy, int64_t K, const float* bias, void* out,
                          int64_t ldo, int k_splits, int64_t split_stride, void* stream);
MMD_API int mmd_gemm_splits(int64_t K, int k_splits);

/* ---------------------------------------------------------------------------------------------------------------
 * Building-block kernels (exported so that tests can check each against the oracle).
 * ------------------------------------------------------------------------------------------------------------- */
/* Frame ingest: decoded BGR uint8 frames [T, in_h, in_w, 3] -> RGB uint8 [T, 3, res, res], aspect-preserving
 * 8-bit INTER_LINEAR resize (bit-exact with cv2.resize: 11-bit fixed-point separable filter) of the longer side to `res`,
 * centred zero padding, BGR->RGB, HWC->CHW.  Replaces the cv2.resize / copyMakeBorder / cvtColor / transpose chain of
 * test/datasets.py:50-72 and demo/liveinfer.py:32-54 (video decoding itself stays on the host).  res % 4 == 0. */
MMD_API int mmd_frame_ingest(const void* frames_bgr_hwc, int n_frames, int in_h, int in_w, void* out_rgb_chw, int res,
                             void* stream);
/* Patch im2col (+ optional (x/255-0.5)/0.5): pixels [T,3,img,img] (u8/bf16/f32) -> A bf16 [T*G*G, k_pad].
 * SiglipVisionEmbeddings Conv2d (TF:models/siglip/modeling_siglip.py:124-130); models/vision_live.py:13. */
MMD_API int mmd_im2col(const void* pixels, int px_dtype, int normalize, void* A, int T, int img, int patch, int k_pad,
                       void* stream);
/* LayerNorm over fp32 rows -> bf16 (or fp32) rows; SiglipEncoderLayer layer_norm1/2, post_layernorm. */
MMD_API int mmd_layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_f32, int64_t rows,
                          int D, float eps, void* stream);
/* Fused SigLIP attention on packed qkv bf16 [T*S, 3*H*dh] -> out bf16 [T*S, H*dh], or with split_hi_lo
 * [T*S, 2*H*dh] = [bf16(o) | bf16(o - bf16(o))] (TF:models/siglip/modeling_siglip.py:252-330). */
MMD_API int mmd_vit_attention(const void* qkv, void* out, int T, int S, int H, int dh, int split_hi_lo, void* stream);
/* resid += sum of split-K planes; out = RMSNorm(resid) * w (bf16 and/or fp32) — Qwen2DecoderLayer residual adds and
 * Qwen2RMSNorm (TF:models/qwen2/modeling_qwen2.py:249-310).  w == NULL: reduction only. */
MMD_API int mmd_resid_add_rmsnorm(float* resid, const float* partial, int n_planes, int64_t plane_stride, const float* w,
                                  void* out_bf16, float* out_f32, int64_t rows, int H, float eps, void* stream);
/* The same with "precise rows": rows with j = prec_of_row[row] >= 0 take their residual update from prec_partial
 * ([n_prec_planes] planes of [P, H], prec_plane_stride elements apart) instead of `partial`, and also get the normalised
 * row as a bf16 hi+lo pair out_hilo[j] = [hi | lo] (2H wide).  prec_of_row == NULL with out_hilo != NULL: every row (j = row).
 * out_bf16 / out_f32 / prec_partial / out_hilo may be NULL. */
MMD_API int mmd_resid_add_rmsnorm_precise(float* resid, const float* partial, int n_planes, int64_t plane_stride,
                                          const float* w, void* out_bf16, float* out_f32, int64_t rows, int H, float eps,
                                          const int* prec_of_row, const float* prec_partial, int n_prec_planes,
                                          int64_t prec_plane_stride, void* out_hilo, void* stream);
/* Final RMSNorm (model.norm) fused with the informative/relevance heads, evaluated ONLY on the rows that are read:
 * for score row i: r = resid[score_rows[i]] + sum planes; n = w * r * rsqrt(mean r^2 + eps); logits_out[i] = n . head_w[0..3];
 * scores_out[i] = {softmax(inf)[1], softmax(rel)[1]}; for lm row i the bf16 normalised row goes to lm_x[i] (lm_head operand).
 * (video_head_live_llava_qwen.py:152-161; test/inference.py:243-244).  prec_* as above (may be NULL / 0). */
MMD_API int mmd_final_norm_heads(const float* resid, const float* partial, int n_planes, int64_t plane_stride, const float* w,
                                 const int* score_rows, int n_score, const int* lm_rows, int n_lm, const float* head_w,
                                 float* logits_out, float* scores_out, void* lm_x, int H, float eps, const int* prec_of_row,
                                 const float* prec_partial, int n_prec_planes, int64_t prec_plane_stride, void* stream);
/* q/k/v bias + RoPE + KV append into the paged pool (TF:models/qwen2/modeling_qwen2.py:127-146,215-233;
 * TF:cache_utils.py:119-120). */
MMD_API int mmd_qkv_finish(const float* partial, int n_planes, int64_t plane_stride, const float* bias,
                           const float* rope_cos, const float* rope_sin, const int* tok_pos, const int* tok_slot,
                           void* q_out, void* kv_layer, int M, int Hq, int Hkv, int dh, void* stream);
/* Chunked-prefill attention over the paged KV pool (see mmd_step for stream_desc / block_tables).
 * o_part: fp32 [n_splits, total_q*Hq, dh]; ml_part: fp32 [n_splits, total_q*Hq, 2]; out bf16 [total_q, Hq*dh]. */
MMD_API int mmd_kv_attention(mmd_ctx*, const void* q, const void* kv_layer, const int* stream_desc, const int* block_tables,
                             int n_streams, int max_n_q, int total_q, int max_kv_len, float* o_part, float* ml_part,
                             void* out, int Hq, int Hkv, int dh, int n_splits /* 0 = auto */, void* stream);
MMD_API int mmd_kv_attention_splits(mmd_ctx*, int max_n_q, int Hq, int Hkv, int n_streams, int max_kv_len);
/* Tap pooling (bilinear / average weights or max) — video_head_live_llava_qwen.py:100-119, vision_live.py:19-25. */
MMD_API int mmd_tap_pool(const void* in, int in_dtype, void* out, int out_dtype, const int* tap_idx, const float* tap_w,
                         int T, int n_in, int n_out, int max_taps, int D, int maxpool, void* stream);
/* informative/relevance heads + sigmoid score on selected rows (video_head_live_llava_qwen.py:160-161;
 * test/inference.py:243-244).  head_w fp32 [4,H] = {inf0, inf1, rel0, rel1}. */
MMD_API int mmd_heads(const float* hidden_f32, const int* rows, const float* head_w, float* logits_out, float* scores_out,
                      int n_rows, int H, void* stream);
/* SigLIP attention-pooling head (the CLS token of models/vision_live.py:26-30 = vision_outputs.pooler_output): one probe query
 * per frame over the S patch tokens.  q fp32 [H*dh]: the projected probe times dh^-0.5 (a constant of the weights);
 * kv bf16 [T*S, 2*H*dh] = [K | V] of the head's in-projection; out fp32 [T, H*dh] (before the head's out_proj). */
MMD_API int mmd_probe_attention(const float* q, const void* kv, float* out, int T, int S, int H, int dh, void* stream);
/* Greedy token pick with the HF repetition penalty (models/modeling_live.py:51-77). */
MMD_API int mmd_argmax(const float* logits, int64_t V, const int64_t* penal_ids, int n_penal, float penalty,
                       int64_t* out_id, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * SigLIP tower (a1/a2): patch embed -> n_layers x {LN, QKV, attention, out-proj, LN, fc1+GELU(tanh), fc2} with an fp32
 * residual stream.  Output: the fp32 residual [T*S, dim] BEFORE post_layernorm (llava path,
 * video_head_live_llava_qwen.py:96-98); the legacy entry (models/vision_live.py:11-31) applies post_ln + pooling on top.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* ln1_w; const float* ln1_b;
  const void* qkv_w;  const float* qkv_b;   /* bf16 [3*dim, dim] (q;k;v stacked), fp32 [3*dim] */
  const void* out_w;  const float* out_b;   /* bf16 [dim, dim], or [dim, 2*dim] = [W | W] when attn_out_split */
  const float* ln2_w; const float* ln2_b;
  const void* fc1_w;  const float* fc1_b;   /* bf16 [mlp, dim] */
  const void* fc2_w;  const float* fc2_b;   /* bf16 [dim, mlp] */
} mmd_vit_layer;

typedef struct {
  int image_size, patch_size, dim, heads, mlp, n_layers, k_pad;
  int attn_out_split;               /* 1: attention output kept as bf16 hi+lo pairs for the out-projection (+8% FLOPs) */
  const void* patch_w;              /* bf16 [dim, k_pad]: Conv2d weight flattened (c, py, px), zero padded */
  const float* patch_b;             /* fp32 [dim] */
  const float* pos_emb;             /* fp32 [S, dim] */
  const mmd_vit_layer* layers;      /* HOST array of n_layers entries */
} mmd_vit_weights;

MMD_API int64_t mmd_vit_workspace_bytes(const mmd_vit_weights*, int T);
MMD_API int mmd_vit_forward(mmd_ctx*, const mmd_vit_weights*, const void* pixels, int px_dtype, int normalize, int T,
                            float* resid_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * mm_projector + GELU + spatial pooling (a2).  video_head_live_llava_qwen.py:90-91 (connector), :100-119
 * (post_projector_pooling).  Two launch sequences:
 *   pooled (pool_group = 4 or 16, linear pooling = bilinear / average, hilo): the source tokens are gathered tap-major
 *     (pool_group consecutive rows = the taps of one output token, padded with weight-0 taps), Linear1 + erf-GELU runs with the
 *     POOLING IN ITS EPILOGUE (MMD_EPI_BF16_HILO_POOL: weighted sum over the group's TMEM lanes, fp32, written as a hi+lo
 *     pair), and Linear2 — which commutes with the linear taps (their weights sum to 1, so the bias is preserved) — runs on
 *     the n_out pooled rows per frame only (49 instead of 169) and writes the frame tokens directly: gather + 2 GEMMs.
 *   generic (pool_group = 0; max pooling does not commute with Linear2): gather the distinct source tokens, Linear1 +
 *     GELU, Linear2 with an fp32 epilogue, tap pooling kernel.
 * Either way the frame tokens are rounded to bf16 exactly once, at the output.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int vit_dim, hidden, n_src_tokens /* S */, n_gather, n_out, max_taps, maxpool;
  int hilo;                          /* 1: operands as bf16 hi+lo pairs, weights stored as [W | W] (K doubled) */
  const void* w1; const float* b1;   /* bf16 [hidden, vit_dim]  (or [hidden, 2*vit_dim] when hilo) */
  const void* w2; const float* b2;   /* bf16 [hidden, hidden]   (or [hidden, 2*hidden]  when hilo) */
  const int* gather_idx;             /* int32 [n_gather]: source token of each gathered row */
  const int* tap_idx;                /* int32 [n_out, max_taps]: index INTO THE GATHERED set, -1 = end */
  const float* tap_w;                /* fp32 [n_out, max_taps] */
  int pool_group;                    /* 0 = generic path; 4 / 16 = taps per output token of the pooled path */
  const int* pool_gather_idx;        /* int32 [n_out * pool_group]: source token of tap slot (o, i) */
  const float* pool_row_w;           /* fp32 [n_out * pool_group]: its weight (0 for padding slots) */
} mmd_projector_weights;

MMD_API int64_t mmd_projector_workspace_bytes(const mmd_projector_weights*, int T);
/* out: [T*n_out, hidden] in out_dtype (MMD_DT_BF16 = the model dtype the reference returns, or MMD_DT_F32 = the same
 * values before the final rounding). */
MMD_API int mmd_projector_pool(mmd_ctx*, const mmd_projector_weights*, const float* vit_resid, int T, void* out,
                               int out_dtype, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Decoder step (a4/a5/a7): embed/concat -> n_layers x {RMSNorm, QKV(+bias,RoPE,KV append), KV-append attention, o_proj,
 * RMSNorm, SwiGLU MLP} -> final RMSNorm -> informative/relevance heads (+ lm_head on requested rows only).
 * VideoHeadLiveLlavaQwenForCausalLM.forward (video_head_live_llava_qwen.py:121-205) as driven by
 * LiveInferForBenchmark._encode_frame/_encode_query (test/inference.py:221-255).  Several streams (videos) and several
 * frames per stream can share one step (varlen), weights are read once per step.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* ln1_w;                /* input_layernorm fp32 [H] */
  const void* qkv_w; const float* qkv_b; /* bf16 [(Hq+2Hkv)*dh, H] (q;k;v stacked), fp32 */
  const void* o_w;                   /* bf16 [H, Hq*dh] */
  const float* ln2_w;                /* post_attention_layernorm */
  const void* gate_up_w;             /* bf16 [2*mlp, H], rows interleaved: row 2j = gate_proj row j, row 2j+1 = up_proj row j */
  const void* down_w;                /* bf16 [H, mlp] */
} mmd_dec_layer;

typedef struct {
  int hidden, n_layers, q_heads, kv_heads, head_dim, mlp, vocab, max_pos;
  float rms_eps;
  const mmd_dec_layer* layers;       /* HOST array */
  const float* final_norm_w;         /* fp32 [H] */
  const void* embed;                 /* bf16 [vocab, H] */
  const void* lm_head;               /* bf16 [vocab, H] (may be NULL when no lm rows are ever requested) */
  const float* heads_w;              /* fp32 [4, H] */
  const float* rope_cos; const float* rope_sin; /* fp32 [max_pos, dh/2] */
} mmd_dec_weights;

typedef struct {
  void* pool;                        /* bf16 [n_layers][n_pages][2][kv_heads][MMD_PAGE_TOKENS][head_dim] */
  int64_t layer_stride;              /* elements between layers */
  int n_pages;
} mmd_kv_pool;

typedef struct {
  int n_tokens;                      /* rows of this step (all streams packed) */
  const int* src_row;                /* int32 [n_tokens]: >= 0 embedding id; < 0 row -(v+1) of frame_tokens */
  const void* frame_tokens;          /* bf16 [*, H] (output of mmd_projector_pool) */
  const int* tok_pos;                /* int32 [n_tokens] RoPE position */
  const int* tok_slot;               /* int32 [n_tokens] physical KV slot = page*MMD_PAGE_TOKENS + offset */
  int n_streams;
  const int* stream_desc;            /* int32 [n_streams,4] {q_start, n_q, kv_len(after append), table_off} */
  const int* block_tables;           /* int32 page ids */
  int max_n_q, max_kv_len;           /* host copies of the maxima over streams (grid sizing) */
  int n_score_rows; const int* score_rows;   /* rows at which the heads are evaluated */
  float* head_logits_out;            /* fp32 [n_score_rows,4] */
  float* scores_out;                 /* fp32 [n_score_rows,2] {informative_score, relevance_score} */
  int n_lm_rows; const int* lm_rows; /* rows that need lm_head (query / generation steps), 0 on frame steps */
  float* lm_logits_out;              /* fp32 [n_lm_rows, vocab] */
  /* "precise rows" of passes with more than 128 tokens: the (<= 128) rows whose outputs are read (score / lm rows) get a
   * second gate/up + down pass per layer on bf16 hi+lo operands.  0 rows = off.  Passes of <= 128 tokens carry EVERY row as
   * hi+lo inside the main kernels and ignore these fields. */
  int n_prec_rows; const int* prec_rows;     /* int32 [n_prec_rows] row indices, ascending */
  const int* prec_of_row;            /* int32 [n_tokens]: index into prec_rows, or -1 */
  /* Layer-pipeline stages (one video's decoder split by layers over several GPUs; the weights struct of a stage lists its
   * own layers only): resid_in != NULL starts the pass from this fp32 residual stream [n_tokens, H] instead of the
   * embedding lookup (src_row / frame_tokens unused); resid_out != NULL writes the residual stream after the stage's last
   * layer there and skips model.norm, the heads and lm_head (score / lm rows still select the precise rows). */
  const float* resid_in;
  float* resid_out;
} mmd_step;

MMD_API int64_t mmd_decoder_workspace_bytes(mmd_ctx*, const mmd_dec_weights*, int max_tokens, int max_lm_rows);
MMD_API int mmd_decoder_step(mmd_ctx*, const mmd_dec_weights*, const mmd_kv_pool*, const mmd_step*, void* workspace,
                             int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Grounding-score post-processing of the reference's evaluator: for every video v and smoothing window w
 * (test/evaluate.py:374-392): s = smooth_pred_list(scores[v], w) (:166-167, np.mean with numpy's pairwise summation),
 * p = normalize_pred_list(s) (:170-173), and for every threshold t: pred = p >= t, counts[w][v][t] = {|pred & gold|,
 * |pred | gold|} (calculate_iou, :129-137; IoU = counts[0] / counts[1], 0 when the union is empty).  float64 throughout,
 * bit-identical to the evaluator.  scores [n_videos, t_max] fp64 (rows padded), gold [n_videos, t_max] u8 (is_time_in_span
 * of each frame time), lens [n_videos]; norm_out (optional) [n_windows, n_videos, t_max] receives p; degenerate
 * [n_windows, n_videos] is set to 1 where the list is empty (the evaluator raises there); a constant list gives p = nan and
 * empty predictions, exactly as the evaluator's np.float64 arithmetic does.  n_thresholds <= 64.
 * ------------------------------------------------------------------------------------------------------------- */
MMD_API int mmd_grounding_sweep(const double* scores, const unsigned char* gold, const int* lens, int n_videos, int t_max,
                                const int* windows, int n_windows, const double* thresholds, int n_thresholds, int* counts,
                                double* norm_out, int* degenerate, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): kernels launched by the stage functions so far, and optional CUDA-event timing of
 * named launch sites ("all" or a comma-separated list of mmd_profile_tag_name values) on the launching stream.
 * mmd_profile_stop synchronises on the recorded events and fills ms_sum[tag] / counts[tag] (mmd_profile_num_tags long).
 * ------------------------------------------------------------------------------------------------------------- */
MMD_API unsigned long long mmd_launch_count(mmd_ctx*);
MMD_API int mmd_profile_num_tags(void);
MMD_API const char* mmd_profile_tag_name(int i);
MMD_API int mmd_profile_start(mmd_ctx*, const char* tags_csv);
MMD_API int mmd_profile_stop(mmd_ctx*, float* ms_sum, int* counts);

#ifdef __cplusplus
}
#endif
#endif /* MMDUET_B200_H_ */
