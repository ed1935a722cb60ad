/* molly_b200 -- C ABI of the B200-native omics-embedding path (encode -> project -> merge).
 *
 * This is the drop-in boundary for the ONE hot path of SeedLLM/molly:
 *     OmicsOne.process_omic_sequences        (reference src/model/omics_one.py:49-136)
 *       -> _inject_omic                      (reference src/model/omics_one.py:57-97)
 *          -> EsmForMaskedLM forward         (transformers models/esm/modeling_esm.py, "HF" below)
 *          -> nn.Linear projector            (reference src/model/omics_one.py:22-30, :91)
 *          -> slice-assign into hidden_states(reference src/model/omics_one.py:93-97)
 *
 * Plain C: pointers + sizes, no torch types.  All pointers named *_dev are device pointers valid on the current CUDA
 * device; `stream` is a cudaStream_t passed as void*.  Every entry point returns 0 on success, non-zero otherwise with
 * a message available from molly_last_error().  Nothing here falls back to the CPU.
 */
#ifndef MOLLY_B200_H_
#define MOLLY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define MOLLY_ABI_VERSION 1

enum molly_status {
    MOLLY_STATUS_OK = 0,
    MOLLY_STATUS_INVALID = 1,      /* bad argument (shape / alignment / NULL) */
    MOLLY_STATUS_CUDA = 2,         /* a CUDA runtime / driver call failed */
    MOLLY_STATUS_UNSUPPORTED = 3,  /* shape outside what the sm_100a kernels are instantiated for */
    MOLLY_STATUS_WORKSPACE = 4     /* workspace too small */
};

enum molly_dtype { MOLLY_DTYPE_BF16 = 0, MOLLY_DTYPE_F32 = 1 };
enum molly_position_type { MOLLY_POS_ROTARY = 0, MOLLY_POS_ABSOLUTE = 1 };
enum molly_ffn_type { MOLLY_FFN_GELU = 0, MOLLY_FFN_GLU = 1 };

/* bits of the device-side error flag (int32, OR-ed by kernels; checked lazily by the host wrapper) */
enum molly_err_bits {
    MOLLY_ERRBIT_OOV = 1,        /* token id outside [0, vocab)           -> reference AssertionError, omics_one.py:71-72 */
    MOLLY_ERRBIT_OVERFLOW = 2,   /* start+1+k > T or bad batch index       -> reference RuntimeError (slice shape mismatch) */
    MOLLY_ERRBIT_POSITION = 4,   /* absolute position id >= max_positions  -> reference IndexError inside the encoder */
    MOLLY_ERRBIT_LAYOUT = 8,     /* placeholder runs in input_ids do not match the omic_ids slots (fused-input path only) */
    MOLLY_ERRBIT_TOKEN = 16      /* input_ids entry outside the LLM embedding table -> reference IndexError (nn.Embedding) */
};

/* Mirrors the fields of transformers.EsmConfig the reference reads (HF:161-186, 285-316) + the two OmicsOne fields
 * (omics_one.py:14-30). */
typedef struct molly_encoder_config {
    int32_t hidden_size;
    int32_t num_layers;
    int32_t num_heads;
    int32_t intermediate_size;
    int32_t vocab_size;
    int32_t pad_token_id;          /* embedding padding_idx / position offset (HF:168, 971-984)            */
    int32_t mask_token_id;         /* token-dropout target (HF:213-222)                                     */
    int32_t position_type;         /* molly_position_type                                                   */
    int32_t max_positions;
    int32_t ffn_type;              /* molly_ffn_type                                                        */
    int32_t token_dropout;         /* 0/1                                                                   */
    int32_t emb_layer_norm_before; /* 0/1                                                                   */
    float layer_norm_eps;
    int32_t llm_hidden_size;       /* D: text_config.hidden_size                                            */
    int32_t project_token_num;     /* K cap: config.{dna_rna,protein}_project_token_num (omics_one.py:25,30) */
} molly_encoder_config;

/* Device pointers to weights already packed by the host (molly_b200/packing.py).  Matrices are bf16 row-major
 * [out_features, in_features] exactly like nn.Linear.weight; vectors are fp32.  Per-layer members are HOST arrays of
 * num_layers device pointers.  The library does not take ownership. */
typedef struct molly_encoder_weights {
    const void* word_emb_dev;            /* bf16 [vocab, h]         esm.embeddings.word_embeddings.weight */
    const void* pos_emb_dev;             /* bf16 [max_pos, h]|NULL  esm.embeddings.position_embeddings.weight */
    const float* emb_ln_w_dev;           /* fp32 [h]|NULL           esm.embeddings.layer_norm.* */
    const float* emb_ln_b_dev;
    const float* rope_cos_dev;           /* fp32 [rope_len, d/2]    cos(t * 10000^(-2i/d))  (HF:81-115) */
    const float* rope_sin_dev;
    int32_t rope_len;
    const float* rope_cos_t_dev;         /* fp32 [d/2, rope_len]: the same tables, frequency-major, read by the QKV GEMM */
    const float* rope_sin_t_dev;         /*   epilogue that rotates q,k itself when head_dim <= 64 */
    const float* const* ln1_w_dev;       /* attention.LayerNorm */
    const float* const* ln1_b_dev;
    const void* const* w_qkv_dev;        /* bf16 [3h, h] = cat(Wq, Wk, Wv)  (HF:329-335); q *= d^-1/2 happens in the epilogue */
    const float* const* b_qkv_dev;       /* fp32 [3h]   = cat(bq, bk, bv) */
    const void* const* w_attn_out_dev;   /* bf16 [h, h]  attention.output.dense */
    const float* const* b_attn_out_dev;
    const float* const* ln2_w_dev;       /* LayerNorm (pre-FFN) */
    const float* const* ln2_b_dev;
    const void* const* w_ffn1_dev;       /* GELU: bf16 [F, h]; GLU: bf16 [2F, h] rows interleaved (x1_0, x2_0, x1_1, ...) */
    const float* const* b_ffn1_dev;      /* fp32 [F] | array of NULL for GLU */
    const void* const* w_ffn2_dev;       /* bf16 [h, F] */
    const float* const* b_ffn2_dev;      /* fp32 [h] | array of NULL for GLU */
    const float* final_ln_w_dev;         /* esm.encoder.emb_layer_norm_after */
    const float* final_ln_b_dev;
    const void* w_proj_dev;              /* bf16 [D, h]  {dna_rna,protein}_projector.weight */
    const float* b_proj_dev;             /* fp32 [D] */
} molly_encoder_weights;

typedef struct molly_encoder molly_encoder_t;

/* ---- encoder handle: replaces `self.{dna_rna,protein}_model` + `self.*_projector` of OmicsOne (omics_one.py:18-30) ---- */
int molly_encoder_create(const molly_encoder_config* cfg, const molly_encoder_weights* w, molly_encoder_t** out);
void molly_encoder_destroy(molly_encoder_t* enc);
size_t molly_encoder_workspace_bytes(const molly_encoder_t* enc, int32_t n_seq, int32_t k_tokens);

/* ---- THE hot path: replaces one `_inject_omic(...)` call (omics_one.py:57-97) for one modality --------------------
 * ids_dev       int64 [n_seq, k_tokens]   == torch.stack(omic_ids)                  (:69)
 * seq_table_dev int32 [n_seq, 2]          (b, start_pos) per sequence; start -1 skips  (:93-95)
 * hidden_states_dev  [B, T, D] bf16|fp32, contiguous, mutated IN PLACE at rows b*T + start+1+j, j < min(K cap, k_tokens)
 * err_flag_dev  int32, OR-ed with molly_err_bits (never cleared by the library); may be NULL
 * enc_out_save_dev  optional bf16 [n_seq*k_tokens, h]: post-final-LayerNorm encoder output kept for molly_project_bwd */
int molly_encode_project_merge_fwd(molly_encoder_t* enc, const int64_t* ids_dev, const int32_t* seq_table_dev,
                                   int32_t n_seq, int32_t k_tokens, void* hidden_states_dev, int32_t hs_dtype,
                                   int32_t B, int32_t T, int32_t D, void* workspace_dev, size_t workspace_bytes,
                                   int32_t* err_flag_dev, void* enc_out_save_dev, void* stream);

/* encoder only: `EsmForMaskedLM(ids, attention_mask=ids!=1, output_hidden_states=True).hidden_states[-1]` as bf16
 * [n_seq*k_tokens, h]  (omics_one.py:73-88; also the input of the pooled heads, embed_text.py:112-129) */
int molly_encode_fwd(molly_encoder_t* enc, const int64_t* ids_dev, int32_t n_seq, int32_t k_tokens, void* out_dev,
                     void* workspace_dev, size_t workspace_bytes, int32_t* err_flag_dev, void* stream);

/* masked mean-pool (embed_text.py:112-129) / CLS readout (baselines/model.py:104-120) over an encoder output */
int molly_pool_fwd(const void* enc_out_dev, const int64_t* ids_dev, int32_t n_seq, int32_t k_tokens, int32_t h,
                   int32_t mode /*0 = masked mean, 1 = CLS (token 0)*/, float* out_dev /*fp32 [n_seq, h]*/, void* stream);

/* ---- placeholder scan: positions of <|dna_pad|>/<|rna_pad|>/<|protein_pad|> in input_ids --------------------------
 * Reproduces from `input_ids` alone the index set the reference takes from host metadata
 * (`info["start"]+1+j`, omics_one.py:93-97; layout produced by omics_dataset.py:270-288).
 * out_pos_dev  int32 [B, T]: ascending token positions of the placeholders of sample b (first out_counts[b] valid)
 * out_kind_dev int32 [B, T]: 0 dna, 1 rna, 2 protein for each listed position                                     */
int molly_placeholder_scan(const int64_t* input_ids_dev, int32_t B, int32_t T, const int64_t pad_token_ids[3],
                           int32_t* out_pos_dev, int32_t* out_kind_dev, int32_t* out_counts_dev, void* stream);

/* ---- the input producer on the device -------------------------------------------------------
 * molly_placeholder_runs: per sample the runs of *_pad tokens in text order (first position, kind 0/1/2, length) and
 *   pos_j[b,t] = index of position t inside its run (-1 for non-placeholder tokens and for runs beyond the sample's
 *   n_slots, which the reference's zip never reaches)                                       (omics_dataset.py:270-288)
 * molly_build_seq_table: seq_table[n] = (b, run_start[b][slot]-1): exactly info["start"] of the reference, paired with the
 *   omic_ids slots BY INDEX like the reference's zip (omics_one.py:105); mismatches OR MOLLY_ERRBIT_LAYOUT
 * molly_placeholder_reject: runs that molly_build_seq_table will reject (shorter than the K cap of their kind, kind not
 *   the modality of the omic_ids slot they pair with -- slot_expect[b][r] 0 dna/rna, 1 protein, -1 none) get pos_j = -1
 *   again, so the lookup below embeds them normally: no row of inputs_embeds is ever left unwritten
 * molly_embed_tokens_skip: inputs_embeds = embed_tokens(input_ids) (omics_one.py:164, :209) for every row the omics path
 *   will not overwrite (j >= K cap or not a placeholder) -- the overwritten rows are never read or written            */
int molly_placeholder_runs(const int64_t* input_ids_dev, int32_t B, int32_t T, const int64_t pad_token_ids[3],
                           const int32_t* n_slots_dev /*[B] omic_ids slots per sample, or NULL*/, int32_t max_runs, int32_t* run_start_dev /*[B,max_runs]*/, int32_t* run_kind_dev,
                           int32_t* run_len_dev, int32_t* n_runs_dev /*[B]*/, int32_t* pos_j_dev /*[B,T]*/, void* stream);
int molly_build_seq_table(const int32_t* b_idx_dev, const int32_t* slot_idx_dev, int32_t n, const int32_t* run_start_dev,
                          const int32_t* run_kind_dev, const int32_t* run_len_dev, const int32_t* n_runs_dev,
                          int32_t max_runs, int32_t expect_protein, int32_t k_need, int32_t* seq_table_dev /*[n,2]*/,
                          int32_t* err_flag_dev, void* stream);
int molly_placeholder_reject(int32_t* pos_j_dev /*[B,T]*/, const int32_t* run_start_dev, const int32_t* run_kind_dev,
                             const int32_t* run_len_dev, const int32_t* n_runs_dev, const int32_t* slot_expect_dev /*[B,max_runs]*/,
                             int32_t B, int32_t T, int32_t max_runs, int32_t cap_dna_rna, int32_t cap_protein, void* stream);
int molly_embed_tokens_skip(const int64_t* input_ids_dev, const int32_t* pos_j_dev, const int64_t pad_token_ids[3],
                            int32_t cap_dna_rna, int32_t cap_protein, const void* table_dev /*[vocab,D]*/, int32_t dtype,
                            int32_t vocab, int32_t D, void* out_dev /*[B,T,D]*/, int32_t B, int32_t T,
                            int32_t* err_flag_dev, void* stream);

/* ---- projector backward (training, --train-mlp): grads of `nn.Linear` projector through the slice-assign ----------
 * d_hidden_dev  [B,T,D] grad wrt merged hidden_states (bf16|fp32).  Rows written by the forward are gathered:
 *   dW [D,h] (fp32) = dY^T X ; db [D] (fp32) = sum dY  (overwritten, the caller accumulates);  those rows of d_hidden
 *   are then ZEROED in place so the embedding table receives exactly zero grad there (autograd of the reference's
 *   slice-assign).                                                                                                   */
int molly_project_bwd(molly_encoder_t* enc, void* d_hidden_dev, int32_t hs_dtype, const int32_t* seq_table_dev,
                      int32_t n_seq, int32_t k_tokens, int32_t B, int32_t T, int32_t D, const void* enc_out_save_dev,
                      float* d_weight_dev, float* d_bias_dev, int32_t zero_rows, void* workspace_dev,
                      size_t workspace_bytes, void* stream);

/* ---- single kernels (unit-parity surface; the hot path above is composed of exactly these) ----------------------- */
enum molly_epilogue {
    MOLLY_EPI_BIAS = 0,          /* out = A W^T + b                       (bf16|fp32 out) */
    MOLLY_EPI_BIAS_GELU = 1,     /* out = gelu_erf(A W^T + b)             (HF:57-61, 406-414) */
    MOLLY_EPI_BIAS_RESIDUAL = 2, /* out = A W^T + b + residual, fp32      (HF:365-375, 417-427) */
    MOLLY_EPI_GLU = 3,           /* out[:, j] = silu(acc[:, 2j]) * acc[:, 2j+1]   (NT-v2 gated FFN) */
    MOLLY_EPI_SCATTER = 4,       /* out[b*T + start+1+j, :] = A W^T + b   (omics_one.py:91-97 fused) */
    MOLLY_EPI_BIAS_ROPE = 5      /* BIAS + scale, then NeoX rotary on out[:, :rope_cols] per head (HF:341-344, 45-123) */
};
int molly_gemm_bf16(const void* a_dev, int32_t lda, const void* w_dev, int32_t ldw, int32_t M, int32_t N, int32_t K,
                    int32_t epilogue, const float* bias_dev, const float* residual_dev, void* out_dev,
                    int32_t out_dtype, int32_t ldo, const int32_t* seq_table_dev, int32_t seq_k_tokens, int32_t B,
                    int32_t T, int32_t k_cap, int32_t* err_flag_dev,
                    int32_t scale_cols /*BIAS / BIAS_ROPE: out[:, :scale_cols] *= scale after the bias (q *= d^-1/2, HF:341)*/,
                    float scale, const float* rope_cos_t_dev /*BIAS_ROPE: fp32 [head_dim/2, rope_len]*/,
                    const float* rope_sin_t_dev, int32_t rope_len, int32_t rope_cols, int32_t rope_head_dim, void* stream);
int molly_layernorm(const float* x_dev, const float* w_dev, const float* b_dev, int32_t rows, int32_t h, float eps,
                    void* out_dev, int32_t out_dtype, void* stream);
int molly_embed(const int64_t* ids_dev, int32_t n_seq, int32_t k_tokens, const molly_encoder_config* cfg,
                const void* word_emb_dev, const void* pos_emb_dev, float* x_dev /*fp32 [n_seq*k, h]*/,
                int32_t* kv_info_dev /*[n_seq,2] = (last non-pad index + 1, #non-pad)*/,
                uint8_t* key_mask_dev /*[n_seq*k] 1 = id != 1*/, int32_t* err_flag_dev, void* stream);
int molly_rotary(void* qkv_dev /*bf16 [rows, 3h] in place on q and k*/, int32_t rows, int32_t k_tokens, int32_t h,
                 int32_t heads, const float* cos_dev, const float* sin_dev, void* stream);
int molly_attention(const void* qkv_dev /*bf16 [n_seq*k, 3h]*/, int32_t n_seq, int32_t k_tokens, int32_t h,
                    int32_t heads, const int32_t* kv_info_dev /*[n_seq,2] from molly_embed*/,
                    const uint8_t* key_mask_dev, void* out_dev /*bf16 [n_seq*k, h]*/, void* stream);
/* same, additionally writing the row log-sum-exp of the scores in the log2 domain (P = exp2(S*log2e - lse2)),
 * fp32 [n_seq, heads, k_tokens]: the statistic the attention backward starts from */
int molly_attention_lse(const void* qkv_dev, int32_t n_seq, int32_t k_tokens, int32_t h, int32_t heads,
                        const int32_t* kv_info_dev, const uint8_t* key_mask_dev, void* out_dev, float* lse2_dev,
                        void* stream);
/* Attention backward (first piece of the encoder backward; autograd of HF:257-282 given the packed, scaled and
 * rotated q', k', v of the forward): d_qkv [n_seq*k, 3h] bf16 receives d(q'), d(k'), d(v) from d_out [n_seq*k, h] bf16, the
 * forward output `out` and its lse2.  delta_ws: fp32 scratch of n_seq*heads*k elements. */
int molly_attention_bwd(const void* qkv_dev, const void* out_dev, const void* d_out_dev, const float* lse2_dev, int32_t n_seq,
                        int32_t k_tokens, int32_t h, int32_t heads, const int32_t* kv_info_dev, const uint8_t* key_mask_dev,
                        void* d_qkv_dev, float* delta_ws_dev, void* stream);
/* ---- encoder backward building blocks (--train-bio, src/utils/tools.py:313-331): autograd of the HF modules
 * the forward kernels replace.  All activations bf16 row-major unless noted; gradients of parameters fp32. */
/* d_weight[N,K] = dy[M,N]^T x[M,K], d_bias[N] = colsum(dy) (d_bias_dev may be NULL): autograd of nn.Linear.  The workspace
 * ((N + K) * roundup8(M) * 2 B) is only used by the MOLLY_WGRAD_TRANSPOSE=1 variant */
int molly_linear_wgrad(const void* dy_dev, const void* x_dev, int32_t M, int32_t N, int32_t K, float* d_weight_dev,
                       float* d_bias_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
/* dy[n*k + j, :] = d_hidden[b, start+1+j, :] for j < k_cap (0 beyond): autograd of the slice-assign, omics_one.py:93-97;
 * zero_rows != 0 additionally zeroes those rows of d_hidden (the gradient that flows on to embed_tokens) */
int molly_gather_rows(void* d_hidden_dev, int32_t hs_dtype, const int32_t* seq_table_dev, int32_t n_seq, int32_t k_tokens,
                      int32_t k_cap, int32_t B, int32_t T, int32_t D, void* dy_dev, int32_t zero_rows, void* stream);
int molly_transpose_bf16(const void* in_dev /*[rows, cols]*/, int32_t rows, int32_t cols, void* out_dev /*[cols, rows]*/,
                         void* stream);
/* LayerNorm backward (HF:394, 479, 511): d_x (fp32) = or += dx; stats_dev fp32 [rows, 2] scratch; d_gamma / d_beta (fp32,
 * pre-zeroed, accumulated with atomics) may be NULL */
int molly_layernorm_bwd(const float* x_dev, const void* dy_dev, const float* gamma_dev, int32_t rows, int32_t h, float eps,
                        float* d_x_dev, int32_t accumulate, float* stats_dev, float* d_gamma_dev, float* d_beta_dev,
                        void* stream);
/* FFN activation forward + backward in one pass: glu == 0: act = gelu_erf(pre) [rows, f_out], d_pre likewise (HF:57-61);
 * glu != 0: pre = (a,b) interleaved [rows, 2 f_out], act = silu(a) * b, d_pre interleaved (NT-v2 gated FFN).
 * d_act_dev == NULL: forward only (act is written, d_pre is not touched) */
int molly_act_fwd_bwd(int32_t glu, const void* pre_dev, const void* d_act_dev, int64_t rows, int32_t f_out, void* act_dev,
                      void* d_pre_dev, void* stream);
int molly_cast_f32_bf16(const float* in_dev, int64_t n, void* out_dev, void* stream);
int molly_scale_cols(void* x_dev /*bf16 [rows, ld]*/, int32_t rows, int32_t ld, int32_t cols, float scale, void* stream);
/* table[index[r], :] += scale[r] * src[r, :] (fp32 atomics): autograd of the embedding gathers (HF:189-236) */
int molly_scatter_add_rows(const float* src_dev, const int32_t* index_dev, const float* scale_dev, int32_t rows, int32_t h,
                           float* table_dev, void* stream);
/* ---- the encoder training step, orchestrated natively (--train-bio: src/utils/tools.py:313-331 sets
 * requires_grad on every encoder parameter; the reference then runs HF EsmModel under torch autograd).  One call runs the
 * forward of all layers and keeps what the backward needs in the caller-owned `tape`; one call runs the backward of a range
 * of layers and writes fp32 parameter gradients into `grads_dev`, laid out as molly_encoder_grad_layout() says, in the
 * PACKED layout of the weights (q,k,v concatenated like cat(Wq,Wk,Wv); GLU rows interleaved like w_ffn1).
 * recompute != 0: the tape keeps only each layer's fp32 input and ONE activation slot that the backward refills layer by
 * layer (activation checkpointing); 0: every layer's activations are kept. */
enum molly_grad_slot {            /* per-layer gradient tensors, in layout order: vectors first, then matrices */
    MOLLY_GRAD_LN2_W = 0, MOLLY_GRAD_LN2_B, MOLLY_GRAD_B_FFN2, MOLLY_GRAD_B_FFN1, MOLLY_GRAD_LN1_W, MOLLY_GRAD_LN1_B,
    MOLLY_GRAD_B_O, MOLLY_GRAD_B_QKV, MOLLY_GRAD_W_FFN2, MOLLY_GRAD_W_FFN1, MOLLY_GRAD_W_O, MOLLY_GRAD_W_QKV, MOLLY_GRAD_SLOTS
};
enum molly_grad_tail_slot {       /* after the L layer groups */
    MOLLY_GRAD_TAIL_FINAL_LN_W = 0, MOLLY_GRAD_TAIL_FINAL_LN_B, MOLLY_GRAD_TAIL_WORD_EMB, MOLLY_GRAD_TAIL_POS_EMB,
    MOLLY_GRAD_TAIL_SLOTS
};
int molly_encoder_train_sizes(const molly_encoder_t* enc, int32_t n_seq, int32_t k_tokens, int32_t recompute,
                              size_t* tape_bytes, size_t* workspace_bytes, int64_t* grad_floats);
/* layer_offsets[MOLLY_GRAD_SLOTS]: float offset of each slot inside a layer group (-1: the encoder has no such parameter);
 * layer l's group starts at l * *layer_group_floats; tail_offsets[MOLLY_GRAD_TAIL_SLOTS]: absolute float offsets (-1: none) */
int molly_encoder_grad_layout(const molly_encoder_t* enc, int64_t* layer_offsets, int64_t* layer_group_floats,
                              int64_t* tail_offsets, int64_t* total_floats);
/* out_dev: bf16 [n_seq*k, h] = hidden_states[-1] (numerically the inference forward: same kernels, same order) */
int molly_encode_train_fwd(molly_encoder_t* enc, const int64_t* ids_dev, int32_t n_seq, int32_t k_tokens, void* out_dev,
                           void* tape_dev, size_t tape_bytes, int32_t recompute, int32_t* err_flag_dev, void* stream);
/* Backward of layers layer_begin, layer_begin-1, ..., layer_end (layer index num_layers = emb_layer_norm_after, which needs
 * d_out_dev = d(loss)/d(hidden_states[-1]), bf16 [n_seq*k, h]).  The running gradient of the residual stream lives in the
 * first n_seq*k*h floats of the workspace between calls: after layer 0 it is d(embedding output), the input of
 * molly_scatter_add_rows for the embedding tables.  The word / position embedding slots of grads_dev are not written here.
 * Calls must walk the layers top-down without gaps (num_layers first): each layer's backward zeroes the gradient group of the
 * layer below it and deposits that layer's b_ffn2 gradient there before that layer's own call accumulates into the group. */
int molly_encode_train_bwd(molly_encoder_t* enc, int32_t n_seq, int32_t k_tokens, void* tape_dev, size_t tape_bytes,
                           int32_t recompute, const void* d_out_dev, float* grads_dev, void* workspace_dev,
                           size_t workspace_bytes, int32_t layer_begin, int32_t layer_end, void* stream);
/* bring-up aid: when non-NULL, CTA 0 of the attention kernel records clock64() stamps into timeline_dev
 * (int64 [2 roles][64 iterations][8 slots]); NULL (default) disables it */
int molly_attention_debug(long long* timeline_dev);
int molly_merge_rows(const void* src_dev /*[n_seq*k, D]*/, const int32_t* seq_table_dev, int32_t n_seq,
                     int32_t k_tokens, int32_t k_cap, void* hidden_states_dev, int32_t dtype, int32_t B, int32_t T,
                     int32_t D, int32_t* err_flag_dev, void* stream);

/* ---- per-kernel-family timing with CUDA events on the launching stream (bench.py's roofline numbers) ------------
 * molly_profile_start() arms it; every kernel launched by this library afterwards is bracketed by two events;
 * molly_profile_stop() synchronises them and fills MOLLY_PROFILE_FAMILIES entries.  Off by default (zero overhead). */
#define MOLLY_PROFILE_FAMILIES 14
typedef struct molly_profile_entry {
    const char* name;      /* embed, layernorm, gemm_qkv, rotary, attention, gemm_attn_out, gemm_ffn1, gemm_ffn2, gemm_proj,
                              gemm_other, merge, other, attention_bwd, rowwise_bwd */
    int32_t launches;
    int32_t work_is_flops; /* 1: `work` is algorithmic FLOP; 0: algorithmic HBM bytes */
    double total_ms;
    double work;
} molly_profile_entry;
int molly_profile_start(void);
int molly_profile_stop(molly_profile_entry* out, int32_t max_entries);

const char* molly_last_error(void);
int molly_abi_version(void);
int molly_kernel_launch_count(void); /* kernels launched by this library since load (bench `gpu_launches`) */
/* A host that captured this library's launches into a CUDA graph reports each replay here (n = launches counted while
 * capturing), so that the counter keeps meaning "kernels of this library that ran" */
int molly_add_kernel_launches(int32_t n);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MOLLY_B200_H_ */
