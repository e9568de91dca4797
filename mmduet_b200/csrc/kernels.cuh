// Launchers of the non-GEMM kernels (elementwise.cu, vit_attention.cu, kv_attention.cu).  All return 0 or a negative code.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace mmd {

enum DType : int { DT_U8 = 0, DT_BF16 = 1, DT_F32 = 2 };

int launch_frame_ingest(const uint8_t* frames, int T, int H, int W, uint8_t* out, int res, cudaStream_t s);
int launch_im2col(const void* px, int px_dtype, int normalize, __nv_bfloat16* A, int T, int C, int img, int P, int Kpad,
                  cudaStream_t s);
int launch_broadcast_rows(const float* src, float* dst, long long rows, int S, int D, cudaStream_t s);
int launch_layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_f32, long long rows, int D,
                     float eps, cudaStream_t s);
// x[row] += add_bias + sum_s planes[s][row]  (written back), then LayerNorm -> out; gamma == nullptr: only the update
int launch_resid_add_layernorm(float* x, const float* planes, int n_planes, long long plane_stride, const float* add_bias,
                               const float* gamma, const float* beta, void* out, int out_f32, long long rows, int D, float eps,
                               cudaStream_t s);
int launch_resid_add_rmsnorm(float* resid, const float* partial, int n_planes, long long plane_stride, const float* w,
                             __nv_bfloat16* out_bf16, float* out_f32, long long rows, int H, float eps, cudaStream_t s);
// the same with "precise rows": see elementwise.cu (PreciseRows)
int launch_resid_add_rmsnorm_precise(float* resid, const float* partial, int n_planes, long long plane_stride, const float* w,
                                     __nv_bfloat16* out_bf16, float* out_f32, long long rows, int H, float eps, const int* prec_of_row,
                                     const float* prec_partial, int n_prec_planes, long long prec_plane_stride, __nv_bfloat16* out_hilo,
                                     cudaStream_t s, long long pair_offset = 0, int append_rows = 0, int n_prec = 0);
// SwiGLU of the precise rows from the raw (gate, up) pre-activations of their hi and lo copies -> two appended bf16 rows each
int launch_swiglu_from_raw(const float* raw, long long ld_raw, __nv_bfloat16* h_rows, long long ldh, int P, int I, cudaStream_t s);
// final RMSNorm fused with the informative/relevance heads on the score rows (+ bf16 normalised rows for lm_head)
int launch_final_norm_heads(const float* resid, const float* partial, int n_planes, long long plane_stride, const float* w,
                            const int* score_rows, int n_score, const int* lm_rows, int n_lm, const float* head_w, float* logits_out,
                            float* scores_out, __nv_bfloat16* lm_x, int H, float eps, const int* prec_of_row, const float* prec_partial,
                            int n_prec_planes, long long prec_plane_stride, cudaStream_t s, long long pair_offset = 0);
int launch_qkv_finish(const float* partial, int n_planes, long long plane_stride, const float* bias, const float* cos_tab,
                      const float* sin_tab, const int* tok_pos, const int* tok_slot, __nv_bfloat16* q_out,
                      __nv_bfloat16* kv_layer, int M, int Hq, int Hkv, int dh, int page_tokens, cudaStream_t s,
                      const int* prec_of_row = nullptr, int n_prec = 0);
int launch_gather_rows_bf16_to_f32(const __nv_bfloat16* table, const __nv_bfloat16* other, const int* src_row, float* dst,
                                   long long rows, int H, cudaStream_t s);
int launch_gather_rows_f32_to_bf16(const float* src, const int* idx, __nv_bfloat16* dst, int T, int S, int G, int D, int hilo, cudaStream_t s);
int launch_tap_pool(const void* in, int in_dtype, void* out, int out_dtype, const int* tap_idx, const float* tap_w, int T,
                    int n_in, int n_out, int max_taps, int D, int maxpool, cudaStream_t s);
int launch_probe_attention(const float* q, const __nv_bfloat16* kv, float* out, int T, int S, int H, int dh, cudaStream_t s);
int launch_heads(const float* hidden_f32, const int* rows, const float* head_w, float* logits_out, float* scores_out, int n_rows,
                 int H, cudaStream_t s);
int launch_argmax(const float* partial, int n_planes, long long plane_stride, int V, const long long* penal_ids, int n_penal,
                  float penalty, long long* out_id, float* out_logit, cudaStream_t s);
int launch_splitk_finish_bf16(const float* partial, int n_planes, long long plane_stride, const float* bias,
                              __nv_bfloat16* out, long long rows, int N, int act, cudaStream_t s);

int launch_vit_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int T, int S, int H, int dh, int split_hi_lo, cudaStream_t s);

extern int g_attention_impl;  // 0 = mma.sync attention kernels (default), 1 = tcgen05/TMEM attention kernels
extern int g_kv_decode;       // 1 (default): <= 16 stacked query rows -> HBM-streaming decode kernel (kv_decode_attention.cu)
int kv_attention_pick_splits(int max_rows, int Hkv, int n_streams, int max_kv_len, int num_sms);
int launch_kv_attention(const __nv_bfloat16* q, const __nv_bfloat16* kv_layer, const int* stream_desc, const int* block_tables,
                        int n_streams, int max_n_q, int total_q, int max_kv_len, float* o_part, float* ml_part, __nv_bfloat16* out,
                        int Hq, int Hkv, int dh, int page_tokens, int n_splits, cudaStream_t s);

}  // namespace mmd
