// Internal (C++) launch interface between the translation units; the C ABI in abi.cu wraps these.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace molly {

enum { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RESID = 2, EPI_GLU = 3, EPI_SCATTER = 4, EPI_BIAS_ROPE = 5,
       EPI_BIAS_ACCUM = 6 };   // fp32 out += acc + bias through a TMA reduce-add: the in-place residual update without reading it
enum { DT_BF16 = 0, DT_F32 = 1 };

void count_launch();
int launch_count();
void add_launches(int n);     // kernels replayed from a CUDA graph the host captured around this library's launches

// ---- optional per-launch profiling (CUDA events on the launching stream; off by default) ----
enum ProfFamily {
    PF_EMBED = 0, PF_LAYERNORM, PF_GEMM_QKV, PF_ROTARY, PF_ATTENTION, PF_GEMM_ATTN_OUT, PF_GEMM_FFN1, PF_GEMM_FFN2,
    PF_GEMM_PROJ, PF_GEMM_OTHER, PF_MERGE, PF_OTHER, PF_ATTENTION_BWD, PF_ROWWISE_BWD, PF_COUNT
};
struct ProfScope {            // brackets ONE kernel launch with two events when profiling is on
    cudaStream_t stream;
    int slot;
    ProfScope(int family, double work, cudaStream_t s);
    ~ProfScope();
};
// Work of a launch that only the device knows (attention over kv_len keys): when profiling is on, returns the device slot
// the NEXT ProfScope's record will read its work from at prof_stop (fill it on `stream` before constructing the scope).
double* prof_next_device_work();
// attention work = coef * k_tokens * h * sum_seq kv_len FLOP (forward: coef 4 = QK^T + PV over kv_len keys for all K query rows)
void prof_attention_work(const int32_t* kv_info, int n_seq, int k_tokens, int h, double coef, cudaStream_t stream);
void set_gemm_family(int f);  // tag for the next gemm_launch calls of this thread
int gemm_family();

// ---- gemm.cu ----
enum GemmTile { GEMM_TILE_PAIR_256 = 0, GEMM_TILE_256 = 1, GEMM_TILE_128 = 2 };
GemmTile gemm_pick_tile(int M, int N, int epi);
int gemm_make_map_a(CUtensorMap* ta, const void* a, int lda, int M, int K);
int gemm_make_map_b(CUtensorMap* tb, const void* w, int ldw, int N, int K, int epi);
int gemm_make_map_c(CUtensorMap* tc, void* out, int out_dtype, int ldo, int M, int n_out);
// EPI_BIAS_RESID reads the residual through `tr` (a map shaped like `tc` over the residual tensor); tr == nullptr: through
// `tc`, i.e. the update is in place on the fp32 stream.
int gemm_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* tc, int M, int N, int K, int epi,
                const float* bias, void* out, int out_dtype, int ldo, const int32_t* seq_table, int seq_k, int B, int T,
                int k_cap, int32_t* err_flag, cudaStream_t stream, int scale_cols = 0, float scale = 1.0f,
                const float* rope_cos_t = nullptr, const float* rope_sin_t = nullptr, int rope_len = 0, int rope_cols = 0,
                int rope_head_dim = 0, const CUtensorMap* tr = nullptr);

// Backward GEMMs on the same kernel, operands taken as they lie in memory (no transposes):
//   GEMM_OPND_K_MN   out[M, N] = a[M, K] b[K, N]        b row-major [K, N]            (dgrad: d_in = d_out W)
//   GEMM_OPND_MN_MN  out[M, N] = a[K, M]^T b[K, N]      both row-major with K as rows  (wgrad: dW = dY^T X), fp32 out,
//                    split over K across the SMs when M x N has too few tiles (partials meet in a TMA reduce-add)
enum { GEMM_OPND_KK = 0, GEMM_OPND_K_MN = 1, GEMM_OPND_MN_MN = 2 };
int gemm_make_map_mn(CUtensorMap* t, const void* base, int rows_k, int cols_mn, int ld);
// out_zeroed: the caller has already zeroed `out` (a split-K wgrad accumulates into it)
int gemm_launch_mn(int opnd, const void* a, int lda, const void* b, int ldb, int M, int N, int K, void* out, int out_dtype,
                   int ldo, cudaStream_t stream, bool out_zeroed = false);

// ---- attention.cu ----
void attention_set_debug(long long* buf);
struct AttnMaps {          // views of the packed [rows, 3h] QKV activation
    CUtensorMap q;         // 128-row boxes (Q tiles; K/V tiles of the 128-key kernels)
    CUtensorMap kv64;      // 64-row boxes (K/V tiles of the 64-key kernel)
};
int attention_make_map(AttnMaps* maps, const void* qkv, int rows, int h, int heads);
// lse2 (optional, fp32 [n_seq, heads, k_tokens]): row log-sum-exp of the scores in the log2 domain, for the backward
int attention_launch(const AttnMaps& tqkv, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_len,
                     const uint8_t* key_mask, void* out, cudaStream_t stream, float* lse2 = nullptr);

// ---- attention2.cu ----  one CTA per SM, two 128-row tiles, P in tensor memory (head_dim <= 64)
constexpr int ATTENTION2_DEFAULT = 0;          // MOLLY_ATTN_V2 = 0 | 1 overrides
constexpr int ATTENTION2_POLY_DEFAULT = 0;     // MOLLY_ATTN_POLY = 0 | 1 | 2 overrides
bool attention2_enabled(int head_dim);
int attention2_launch(const AttnMaps& maps, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                      const uint8_t* key_mask, void* out, float* lse2, cudaStream_t stream);

// ---- rowwise.cu ----
struct EmbedArgs {
    int hidden, vocab, pad_id, mask_id, position_type, max_positions, token_dropout;
    int apply_mask;   // 1: x *= (id != 1) inside the gather; 0: caller masks after emb_layer_norm_before
};
int embed_launch(const int64_t* ids, int n_seq, int k_tokens, const EmbedArgs& a, const void* word_emb,
                 const void* pos_emb, float* x, int32_t* kv_len, uint8_t* key_mask, int32_t* err_flag,
                 cudaStream_t stream);
int mask_rows_launch(float* x, const uint8_t* key_mask, int rows, int h, cudaStream_t stream);
int layernorm_launch(const float* x, const float* w, const float* b, int rows, int h, float eps, void* out,
                     int out_dtype, cudaStream_t stream);
// sin_sign = -1: the inverse rotation (rotary backward); q_scale multiplies the rotated q columns (backward of q * d^-1/2)
int rotary_launch(void* qkv, int rows, int k_tokens, int h, int heads, const float* cos_t, const float* sin_t,
                  cudaStream_t stream, float sin_sign = 1.0f, float q_scale = 1.0f);
int pool_launch(const void* enc_out, const int64_t* ids, int n_seq, int k_tokens, int h, int mode, float* out,
                cudaStream_t stream);

// ---- merge.cu ----
int placeholder_scan_launch(const int64_t* input_ids, int B, int T, int64_t pad0, int64_t pad1, int64_t pad2,
                            int32_t* out_pos, int32_t* out_kind, int32_t* out_counts, cudaStream_t stream);
int merge_rows_launch(const void* src, const int32_t* seq_table, int n_seq, int k_tokens, int k_cap, void* hidden,
                      int dtype, int B, int T, int D, int32_t* err_flag, cudaStream_t stream);
int gather_grad_rows_launch(void* d_hidden, int dtype, const int32_t* seq_table, int n_seq, int k_tokens, int k_cap,
                            int B, int T, int D, void* dy_bf16, int zero_rows, cudaStream_t stream);

int placeholder_runs_launch(const int64_t* input_ids, int B, int T, int64_t pad0, int64_t pad1, int64_t pad2,
                            const int32_t* n_slots, int max_runs, int32_t* run_start, int32_t* run_kind, int32_t* run_len, int32_t* n_runs,
                            int32_t* pos_j, cudaStream_t stream);
int placeholder_reject_launch(int32_t* pos_j, const int32_t* run_start, const int32_t* run_kind, const int32_t* run_len,
                              const int32_t* n_runs, const int32_t* slot_expect, int B, int T, int max_runs, int cap_nt,
                              int cap_pr, cudaStream_t stream);
int embed_tokens_skip_launch(const int64_t* input_ids, const int32_t* pos_j, int64_t pad0, int64_t pad1, int cap_nt,
                             int cap_pr, const void* table, int dtype, int vocab, int D, void* out, int B, int T,
                             int32_t* err_flag, cudaStream_t stream);
int build_seq_table_launch(const int32_t* b_idx, const int32_t* slot_idx, int n, const int32_t* run_start,
                           const int32_t* run_kind, const int32_t* run_len, const int32_t* n_runs, int max_runs,
                           int expect_protein, int k_need, int32_t* seq_table, int32_t* err_flag, cudaStream_t stream);

// ---- attention_bwd.cu ----  d(q', k', v) [rows, 3h] bf16 from d(out); delta_ws: fp32 [n_seq * heads * k_tokens] scratch
int attention_bwd_launch(const void* qkv, const void* out, const void* d_out, const float* lse2, int n_seq, int k_tokens, int h,
                         int heads, const int32_t* kv_info, const uint8_t* key_mask, void* d_qkv, float* delta_ws,
                         cudaStream_t stream);

// ---- rowwise_bwd.cu ----
// dy_next_bf16 (optional): bf16 copy of the new d_x; d_bias_next (optional, pre-zeroed): += column sums of dy_next
int ln_bwd_launch(const float* x, const void* dy_bf16, const float* gamma, int rows, int h, float eps, float* d_x,
                  int accumulate, float* stats, float* d_gamma, float* d_beta, cudaStream_t stream,
                  void* dy_next_bf16 = nullptr, float* d_bias_next = nullptr);
// d_bias (optional, GELU backward only, pre-zeroed fp32 [f_out]): += column sums of d_pre
int act_fwd_bwd_launch(int glu, const void* pre, const void* d_act, long long rows, int f_out, void* act, void* d_pre,
                       cudaStream_t stream, float* d_bias = nullptr);
int cast_f32_bf16_launch(const float* in, long long n, void* out, cudaStream_t stream);
int scale_cols_launch(void* x_bf16, int rows, int ld, int cols, float scale, cudaStream_t stream);
int scatter_add_rows_launch(const float* src, const int32_t* index, const float* scale, int rows, int h, float* table,
                            cudaStream_t stream);

// ---- bwd.cu ----
int transpose_bf16_launch(const void* in, int rows, int cols, void* out, cudaStream_t stream);
// zeroed: d_weight and d_bias are already zero (the training step clears a whole layer's gradient group with one memset)
int linear_wgrad_launch(const void* dy_bf16, const void* x_bf16, int M, int N, int K, float* d_weight, float* d_bias,
                        cudaStream_t stream, bool zeroed = false);
int project_bwd_launch(const void* dy_bf16, const void* x_bf16, int M, int D, int h, float* d_weight, float* d_bias,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace molly
