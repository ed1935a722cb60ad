// extern "C" surface declared in include/molly_b200.h.  Orchestrates the per-layer kernel sequence of the encoder
// (HF EsmModel.forward, HF:615-677 -> EsmEncoder.forward, HF:494-514 -> EsmLayer, HF:446-482) on one stream.
#include <math.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/molly_b200.h"
#include <stdlib.h>

#include "common.h"
#include "encoder.h"
#include "kernels.h"

using namespace molly;

namespace {

// MOLLY_RESID_REDUCE=1: the two residual GEMMs of a layer add (acc + bias) into the fp32 stream with TMA reduce-adds instead of
// loading the residual tile, adding and storing it (bit-identical: one fp32 add per element either way)
int residual_epilogue() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOLLY_RESID_REDUCE");
        v = (e != nullptr && e[0] == '1') ? EPI_BIAS_ACCUM : EPI_BIAS_RESID;
    }
    return v;
}

constexpr size_t kAlign = 1024;
size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct Workspace {
    size_t off_x, off_xn, off_qkv, off_attn, off_mid, off_kvinfo, off_mask, total;
};

Workspace layout(const molly_encoder_config& c, int n_seq, int k) {
    const size_t M = static_cast<size_t>(n_seq) * k, h = c.hidden_size, F = c.intermediate_size;
    Workspace w;
    size_t o = 0;
    w.off_x = o;      o += align_up(M * h * 4);
    w.off_xn = o;     o += align_up(M * h * 2);
    w.off_qkv = o;    o += align_up(M * 3 * h * 2);
    w.off_attn = o;   o += align_up(M * h * 2);
    w.off_mid = o;    o += align_up(M * F * 2);
    w.off_kvinfo = o; o += align_up(static_cast<size_t>(n_seq) * 2 * 4);
    w.off_mask = o;   o += align_up(M);
    w.total = o;
    return w;
}

int validate_cfg(const molly_encoder_config& c) {
    MOLLY_CHECK(c.hidden_size > 0 && c.num_layers > 0 && c.num_heads > 0 && c.intermediate_size > 0 && c.vocab_size > 0,
                MOLLY_ERR_INVALID, "encoder config has non-positive sizes");
    MOLLY_CHECK(c.hidden_size % c.num_heads == 0, MOLLY_ERR_INVALID, "hidden_size %d not divisible by heads %d",
                c.hidden_size, c.num_heads);
    const int d = c.hidden_size / c.num_heads;
    MOLLY_CHECK(d == 16 || d == 32 || d == 64 || d == 128, MOLLY_ERR_UNSUPPORTED, "head_dim %d not in {16,32,64,128}", d);
    MOLLY_CHECK(c.hidden_size % 32 == 0 && c.intermediate_size % 32 == 0 && c.llm_hidden_size % 32 == 0,
                MOLLY_ERR_UNSUPPORTED, "hidden / intermediate / llm hidden sizes must be multiples of 32");
    MOLLY_CHECK(c.hidden_size <= 2560, MOLLY_ERR_UNSUPPORTED, "hidden_size %d > 2560 (LayerNorm register tile)", c.hidden_size);
    MOLLY_CHECK(c.position_type == MOLLY_POS_ROTARY || c.position_type == MOLLY_POS_ABSOLUTE, MOLLY_ERR_INVALID,
                "unknown position_type %d", c.position_type);
    MOLLY_CHECK(c.ffn_type == MOLLY_FFN_GELU || c.ffn_type == MOLLY_FFN_GLU, MOLLY_ERR_INVALID, "unknown ffn_type %d",
                c.ffn_type);
    return MOLLY_OK;
}

int build_plan(molly_encoder* e, void* ws, int n_seq, int k, void* final_out) {
    auto& p = e->plan;
    if (p.ws == ws && p.n_seq == n_seq && p.k == k && p.final_out == final_out) return MOLLY_OK;
    const auto& c = e->cfg;
    const Workspace L = layout(c, n_seq, k);
    const int M = n_seq * k, h = c.hidden_size, F = c.intermediate_size;
    auto* base = static_cast<uint8_t*>(ws);
    int rc;
    if ((rc = gemm_make_map_a(&p.tm_xn, base + L.off_xn, h, M, h))) return rc;
    if ((rc = gemm_make_map_a(&p.tm_attn, base + L.off_attn, h, M, h))) return rc;
    if ((rc = gemm_make_map_a(&p.tm_mid, base + L.off_mid, F, M, F))) return rc;
    if ((rc = attention_make_map(&p.tm_qkv, base + L.off_qkv, M, h, c.num_heads))) return rc;
    if ((rc = gemm_make_map_a(&p.tm_final, final_out, h, M, h))) return rc;
    if ((rc = gemm_make_map_c(&p.tc_qkv, base + L.off_qkv, DT_BF16, 3 * h, M, 3 * h))) return rc;
    if ((rc = gemm_make_map_c(&p.tc_x, base + L.off_x, DT_F32, h, M, h))) return rc;
    if ((rc = gemm_make_map_c(&p.tc_mid, base + L.off_mid, DT_BF16, F, M, F))) return rc;
    p.ws = ws; p.n_seq = n_seq; p.k = k; p.final_out = final_out;
    return MOLLY_OK;
}

// Encoder forward up to and including emb_layer_norm_after; result (bf16 [M,h]) lands in `final_out`.
int encode(molly_encoder* e, const int64_t* ids, int n_seq, int k, void* final_out, void* ws, size_t ws_bytes,
           int32_t* err_flag, cudaStream_t stream) {
    const auto& c = e->cfg;
    MOLLY_CHECK(ids != nullptr && ws != nullptr && final_out != nullptr, MOLLY_ERR_INVALID, "encode: NULL pointer");
    MOLLY_CHECK(n_seq > 0 && k > 0, MOLLY_ERR_INVALID, "encode: n_seq=%d k_tokens=%d", n_seq, k);
    MOLLY_CHECK((reinterpret_cast<uintptr_t>(ws) & (kAlign - 1)) == 0, MOLLY_ERR_INVALID, "workspace must be 1024-B aligned");
    const Workspace L = layout(c, n_seq, k);
    MOLLY_CHECK(ws_bytes >= L.total, MOLLY_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, L.total);
    if (c.position_type == MOLLY_POS_ROTARY)
        MOLLY_CHECK(e->w.rope_len >= k && e->w.rope_cos_dev && e->w.rope_sin_dev, MOLLY_ERR_INVALID,
                    "rotary tables cover %d positions < k_tokens %d", e->w.rope_len, k);
    int rc = build_plan(e, ws, n_seq, k, final_out);
    if (rc) return rc;
    const auto& p = e->plan;
    const int M = n_seq * k, h = c.hidden_size, F = c.intermediate_size;
    auto* base = static_cast<uint8_t*>(ws);
    float* x = reinterpret_cast<float*>(base + L.off_x);
    void* xn = base + L.off_xn;
    void* qkv = base + L.off_qkv;
    void* attn = base + L.off_attn;
    void* mid = base + L.off_mid;
    int32_t* kv_info = reinterpret_cast<int32_t*>(base + L.off_kvinfo);
    uint8_t* key_mask = base + L.off_mask;

    EmbedArgs ea{h, c.vocab_size, c.pad_token_id, c.mask_token_id, c.position_type, c.max_positions, c.token_dropout,
                 c.emb_layer_norm_before ? 0 : 1};
    if ((rc = embed_launch(ids, n_seq, k, ea, e->w.word_emb_dev, e->w.pos_emb_dev, x, kv_info, key_mask, err_flag, stream)))
        return rc;
    if (c.emb_layer_norm_before) {     // HF:229-233: LayerNorm, then multiply by the attention mask
        if ((rc = layernorm_launch(x, e->w.emb_ln_w_dev, e->w.emb_ln_b_dev, M, h, c.layer_norm_eps, x, DT_F32, stream)))
            return rc;
        if ((rc = mask_rows_launch(x, key_mask, M, h, stream))) return rc;
    }
    for (int l = 0; l < c.num_layers; ++l) {
        // --- attention block: x = x + Wo * Attn(LN(x)) + bo   (HF:386-403)
        if ((rc = layernorm_launch(x, e->ln1_w[l], e->ln1_b[l], M, h, c.layer_norm_eps, xn, DT_BF16, stream))) return rc;
        // q, k, v = Linear(LN(x)); q *= d^-1/2 BEFORE rotary (HF:329-341) -- folded into the epilogue of one fused GEMM
        set_gemm_family(PF_GEMM_QKV);
        // and, for head_dim <= 64, so is the rotary embedding (HF:343-344): no separate pass over q and k
        const int d = h / c.num_heads;
        const bool rope = c.position_type == MOLLY_POS_ROTARY;
        const bool rope_fused = rope && d <= 64 && e->w.rope_cos_t_dev != nullptr && e->w.rope_sin_t_dev != nullptr;
        if ((rc = gemm_launch(p.tm_xn, e->tm_wqkv[l], &p.tc_qkv, M, 3 * h, h, rope_fused ? EPI_BIAS_ROPE : EPI_BIAS,
                              e->b_qkv[l], qkv, DT_BF16, 3 * h, nullptr, k, 0, 0, 0, nullptr, stream, h, e->q_scale,
                              e->w.rope_cos_t_dev, e->w.rope_sin_t_dev, e->w.rope_len, 2 * h, d)))
            return rc;
        if (rope && !rope_fused)
            if ((rc = rotary_launch(qkv, M, k, h, c.num_heads, e->w.rope_cos_dev, e->w.rope_sin_dev, stream))) return rc;
        if ((rc = attention_launch(p.tm_qkv, n_seq, k, h, c.num_heads, kv_info, key_mask, attn, stream))) return rc;
        set_gemm_family(PF_GEMM_ATTN_OUT);
        if ((rc = gemm_launch(p.tm_attn, e->tm_wo[l], &p.tc_x, M, h, h, residual_epilogue(), e->b_o[l], x, DT_F32, h, nullptr, 0,
                              0, 0, 0, nullptr, stream)))
            return rc;
        // --- feed-forward block: x = x + W2 * act(W1 * LN(x) + b1) + b2   (HF:478-482)
        if ((rc = layernorm_launch(x, e->ln2_w[l], e->ln2_b[l], M, h, c.layer_norm_eps, xn, DT_BF16, stream))) return rc;
        const int epi1 = c.ffn_type == MOLLY_FFN_GLU ? EPI_GLU : EPI_BIAS_GELU;
        set_gemm_family(PF_GEMM_FFN1);
        if ((rc = gemm_launch(p.tm_xn, e->tm_w1[l], &p.tc_mid, M, e->ffn1_n, h, epi1, e->b_ffn1[l], mid, DT_BF16, F,
                              nullptr, 0, 0, 0, 0, nullptr, stream)))
            return rc;
        set_gemm_family(PF_GEMM_FFN2);
        if ((rc = gemm_launch(p.tm_mid, e->tm_w2[l], &p.tc_x, M, h, F, residual_epilogue(), e->b_ffn2[l], x, DT_F32, h, nullptr,
                              0, 0, 0, 0, nullptr, stream)))
            return rc;
    }
    set_gemm_family(PF_GEMM_OTHER);
    // emb_layer_norm_after (HF:511-512) -> hidden_states[-1] (omics_one.py:91)
    return layernorm_launch(x, e->w.final_ln_w_dev, e->w.final_ln_b_dev, M, h, c.layer_norm_eps, final_out, DT_BF16, stream);
}

}  // namespace

namespace {
bool wgrad_through_transposes() {     // MOLLY_WGRAD_TRANSPOSE=1: the first version (explicit bf16 transposes + K-major GEMM)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOLLY_WGRAD_TRANSPOSE");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
}  // namespace

extern "C" {

const char* molly_last_error(void) { return get_last_error(); }
int molly_abi_version(void) { return MOLLY_ABI_VERSION; }
int molly_kernel_launch_count(void) { return launch_count(); }
int molly_add_kernel_launches(int32_t n) {
    MOLLY_CHECK(n >= 0, MOLLY_ERR_INVALID, "molly_add_kernel_launches: n=%d", n);
    add_launches(n);
    return MOLLY_OK;
}

int molly_profile_start(void) { prof_start(); return MOLLY_OK; }
int molly_profile_stop(molly_profile_entry* out, int32_t max_entries) {
    static const char* names[PF_COUNT] = {"embed", "layernorm", "gemm_qkv", "rotary", "attention", "gemm_attn_out",
                                          "gemm_ffn1", "gemm_ffn2", "gemm_proj", "gemm_other", "merge", "other",
                                          "attention_bwd", "rowwise_bwd"};
    static const int is_flops[PF_COUNT] = {0, 0, 1, 0, 1, 1, 1, 1, 1, 1, 0, 0, 1, 0};
    int launches[PF_COUNT]; double ms[PF_COUNT], work[PF_COUNT];
    int rc = prof_stop(launches, ms, work);
    MOLLY_CHECK(out != nullptr && max_entries >= PF_COUNT, MOLLY_ERR_INVALID, "molly_profile_stop: need %d entries", PF_COUNT);
    for (int f = 0; f < PF_COUNT; ++f) {
        out[f].name = names[f]; out[f].launches = launches[f]; out[f].total_ms = ms[f]; out[f].work = work[f];
        out[f].work_is_flops = is_flops[f];
    }
    MOLLY_CHECK(rc == 0, MOLLY_ERR_CUDA, "molly_profile_stop: event timing failed");
    return PF_COUNT == MOLLY_PROFILE_FAMILIES ? MOLLY_OK : MOLLY_ERR_INVALID;
}

int molly_encoder_create(const molly_encoder_config* cfg, const molly_encoder_weights* w, molly_encoder_t** out) {
    MOLLY_CHECK(cfg && w && out, MOLLY_ERR_INVALID, "molly_encoder_create: NULL argument");
    int rc = validate_cfg(*cfg);
    if (rc) return rc;
    const int L = cfg->num_layers, h = cfg->hidden_size, F = cfg->intermediate_size, D = cfg->llm_hidden_size;
    MOLLY_CHECK(w->word_emb_dev && w->final_ln_w_dev && w->final_ln_b_dev && w->w_proj_dev, MOLLY_ERR_INVALID,
                "molly_encoder_create: missing weights");
    MOLLY_CHECK(w->ln1_w_dev && w->ln1_b_dev && w->w_qkv_dev && w->b_qkv_dev && w->w_attn_out_dev && w->b_attn_out_dev &&
                    w->ln2_w_dev && w->ln2_b_dev && w->w_ffn1_dev && w->b_ffn1_dev && w->w_ffn2_dev && w->b_ffn2_dev,
                MOLLY_ERR_INVALID, "molly_encoder_create: missing per-layer pointer arrays");
    if (cfg->emb_layer_norm_before)
        MOLLY_CHECK(w->emb_ln_w_dev && w->emb_ln_b_dev, MOLLY_ERR_INVALID, "emb_layer_norm_before needs emb_ln weights");
    if (cfg->position_type == MOLLY_POS_ABSOLUTE)
        MOLLY_CHECK(w->pos_emb_dev != nullptr, MOLLY_ERR_INVALID, "absolute positions need pos_emb");
    auto* e = new (std::nothrow) molly_encoder();
    MOLLY_CHECK(e != nullptr, MOLLY_ERR_INVALID, "out of host memory");
    e->cfg = *cfg;
    e->w = *w;
    e->ffn1_n = cfg->ffn_type == MOLLY_FFN_GLU ? 2 * F : F;
    e->q_scale = 1.0f / sqrtf(static_cast<float>(cfg->hidden_size / cfg->num_heads));
    auto copyf = [&](std::vector<const float*>& dst, const float* const* src) { dst.assign(src, src + L); };
    auto copyv = [&](std::vector<const void*>& dst, const void* const* src) { dst.assign(src, src + L); };
    copyf(e->ln1_w, w->ln1_w_dev); copyf(e->ln1_b, w->ln1_b_dev); copyf(e->b_qkv, w->b_qkv_dev);
    copyf(e->b_o, w->b_attn_out_dev); copyf(e->ln2_w, w->ln2_w_dev); copyf(e->ln2_b, w->ln2_b_dev);
    copyf(e->b_ffn1, w->b_ffn1_dev); copyf(e->b_ffn2, w->b_ffn2_dev);
    copyv(e->w_qkv, w->w_qkv_dev); copyv(e->w_o, w->w_attn_out_dev); copyv(e->w_ffn1, w->w_ffn1_dev);
    copyv(e->w_ffn2, w->w_ffn2_dev);
    e->tm_wqkv.resize(L); e->tm_wo.resize(L); e->tm_w1.resize(L); e->tm_w2.resize(L);
    for (int l = 0; l < L && rc == 0; ++l) {
        if (!e->w_qkv[l] || !e->w_o[l] || !e->w_ffn1[l] || !e->w_ffn2[l] || !e->ln1_w[l] || !e->ln1_b[l] || !e->ln2_w[l] ||
            !e->ln2_b[l] || !e->b_qkv[l] || !e->b_o[l]) {
            set_last_error("molly_encoder_create: NULL weight pointer in a layer");
            rc = MOLLY_ERR_INVALID;
            break;
        }
        if (cfg->ffn_type == MOLLY_FFN_GELU && (!e->b_ffn1[l] || !e->b_ffn2[l])) {
            set_last_error("molly_encoder_create: GELU FFN needs biases");
            rc = MOLLY_ERR_INVALID;
            break;
        }
        if ((rc = gemm_make_map_b(&e->tm_wqkv[l], e->w_qkv[l], h, 3 * h, h, EPI_BIAS))) break;
        if ((rc = gemm_make_map_b(&e->tm_wo[l], e->w_o[l], h, h, h, EPI_BIAS_RESID))) break;
        if ((rc = gemm_make_map_b(&e->tm_w1[l], e->w_ffn1[l], h, e->ffn1_n, h,
                                  cfg->ffn_type == MOLLY_FFN_GLU ? EPI_GLU : EPI_BIAS_GELU))) break;
        if ((rc = gemm_make_map_b(&e->tm_w2[l], e->w_ffn2[l], F, h, F, EPI_BIAS_RESID))) break;
    }
    if (rc == 0) rc = gemm_make_map_b(&e->tm_wproj, w->w_proj_dev, h, D, h, EPI_SCATTER);
    if (rc) { delete e; return rc; }
    *out = e;
    return MOLLY_OK;
}

void molly_encoder_destroy(molly_encoder_t* enc) { delete enc; }

size_t molly_encoder_workspace_bytes(const molly_encoder_t* enc, int32_t n_seq, int32_t k_tokens) {
    if (!enc || n_seq <= 0 || k_tokens <= 0) return 0;
    return layout(enc->cfg, n_seq, k_tokens).total;
}

int molly_encode_fwd(molly_encoder_t* enc, const int64_t* ids_dev, int32_t n_seq, int32_t k_tokens, void* out_dev,
                     void* workspace_dev, size_t workspace_bytes, int32_t* err_flag_dev, void* stream) {
    MOLLY_CHECK(enc != nullptr, MOLLY_ERR_INVALID, "molly_encode_fwd: NULL encoder");
    return encode(enc, ids_dev, n_seq, k_tokens, out_dev, workspace_dev, workspace_bytes, err_flag_dev,
                  static_cast<cudaStream_t>(stream));
}

int molly_encode_project_merge_fwd(molly_encoder_t* enc, const int64_t* ids_dev, const int32_t* seq_table_dev,
                                   int32_t n_seq, int32_t k_tokens, void* hidden_states_dev, int32_t hs_dtype,
                                   int32_t B, int32_t T, int32_t D, void* workspace_dev, size_t workspace_bytes,
                                   int32_t* err_flag_dev, void* enc_out_save_dev, void* stream) {
    MOLLY_CHECK(enc != nullptr, MOLLY_ERR_INVALID, "molly_encode_project_merge_fwd: NULL encoder");
    MOLLY_CHECK(seq_table_dev && hidden_states_dev, MOLLY_ERR_INVALID, "molly_encode_project_merge_fwd: NULL pointer");
    MOLLY_CHECK(D == enc->cfg.llm_hidden_size, MOLLY_ERR_INVALID, "hidden_states D=%d != projector out_features %d", D,
                enc->cfg.llm_hidden_size);
    MOLLY_CHECK(hs_dtype == MOLLY_DTYPE_BF16 || hs_dtype == MOLLY_DTYPE_F32, MOLLY_ERR_INVALID, "bad hs_dtype %d", hs_dtype);
    MOLLY_CHECK(B > 0 && T > 0, MOLLY_ERR_INVALID, "B=%d T=%d", B, T);
    auto s = static_cast<cudaStream_t>(stream);
    const Workspace L = layout(enc->cfg, n_seq > 0 ? n_seq : 1, k_tokens > 0 ? k_tokens : 1);
    void* final_out = enc_out_save_dev ? enc_out_save_dev : static_cast<uint8_t*>(workspace_dev) + L.off_xn;
    int rc = encode(enc, ids_dev, n_seq, k_tokens, final_out, workspace_dev, workspace_bytes, err_flag_dev, s);
    if (rc) return rc;
    // projector + merge: hidden[b, start+1+j, :] = LN_out[n*K+j, :] Wp^T + bp  for j < min(K cap, K)  (omics_one.py:91-97)
    const int k_cap = enc->cfg.project_token_num < k_tokens ? enc->cfg.project_token_num : k_tokens;
    set_gemm_family(PF_GEMM_PROJ);
    rc = gemm_launch(enc->plan.tm_final, enc->tm_wproj, nullptr, n_seq * k_tokens, D, enc->cfg.hidden_size, EPI_SCATTER,
                     enc->w.b_proj_dev, hidden_states_dev, hs_dtype, D, seq_table_dev, k_tokens, B, T, k_cap, err_flag_dev, s);
    set_gemm_family(PF_GEMM_OTHER);
    return rc;
}

int molly_pool_fwd(const void* enc_out_dev, const int64_t* ids_dev, int32_t n_seq, int32_t k_tokens, int32_t h,
                   int32_t mode, float* out_dev, void* stream) {
    MOLLY_CHECK(enc_out_dev && ids_dev && out_dev && n_seq > 0 && k_tokens > 0 && h > 0, MOLLY_ERR_INVALID,
                "molly_pool_fwd: bad argument");
    return pool_launch(enc_out_dev, ids_dev, n_seq, k_tokens, h, mode, out_dev, static_cast<cudaStream_t>(stream));
}

int molly_placeholder_scan(const int64_t* input_ids_dev, int32_t B, int32_t T, const int64_t pad_token_ids[3],
                           int32_t* out_pos_dev, int32_t* out_kind_dev, int32_t* out_counts_dev, void* stream) {
    MOLLY_CHECK(input_ids_dev && pad_token_ids && out_pos_dev && out_kind_dev && out_counts_dev, MOLLY_ERR_INVALID,
                "molly_placeholder_scan: NULL pointer");
    return placeholder_scan_launch(input_ids_dev, B, T, pad_token_ids[0], pad_token_ids[1], pad_token_ids[2], out_pos_dev,
                                   out_kind_dev, out_counts_dev, static_cast<cudaStream_t>(stream));
}

int molly_project_bwd(molly_encoder_t* enc, void* d_hidden_dev, int32_t hs_dtype, const int32_t* seq_table_dev,
                      int32_t n_seq, int32_t k_tokens, int32_t B, int32_t T, int32_t D, const void* enc_out_save_dev,
                      float* d_weight_dev, float* d_bias_dev, int32_t zero_rows, void* workspace_dev,
                      size_t workspace_bytes, void* stream) {
    MOLLY_CHECK(enc && d_hidden_dev && seq_table_dev && enc_out_save_dev && d_weight_dev && d_bias_dev && workspace_dev,
                MOLLY_ERR_INVALID, "molly_project_bwd: NULL pointer");
    MOLLY_CHECK(D == enc->cfg.llm_hidden_size && n_seq > 0 && k_tokens > 0, MOLLY_ERR_INVALID, "molly_project_bwd: bad shape");
    auto s = static_cast<cudaStream_t>(stream);
    const int M = n_seq * k_tokens, h = enc->cfg.hidden_size;
    const size_t dy_bytes = align_up(static_cast<size_t>(M) * D * 2);
    MOLLY_CHECK(workspace_bytes > dy_bytes, MOLLY_ERR_WORKSPACE, "molly_project_bwd: workspace too small");
    const int k_cap = enc->cfg.project_token_num < k_tokens ? enc->cfg.project_token_num : k_tokens;
    int rc = gather_grad_rows_launch(d_hidden_dev, hs_dtype, seq_table_dev, n_seq, k_tokens, k_cap, B, T, D, workspace_dev,
                                     zero_rows, s);
    if (rc) return rc;
    if (wgrad_through_transposes())
        return project_bwd_launch(workspace_dev, enc_out_save_dev, M, D, h, d_weight_dev, d_bias_dev,
                                  static_cast<uint8_t*>(workspace_dev) + dy_bytes, workspace_bytes - dy_bytes, s);
    return linear_wgrad_launch(workspace_dev, enc_out_save_dev, M, D, h, d_weight_dev, d_bias_dev, s);
}

// ------------------------------------ encoder backward building blocks ------------------------------------
int molly_linear_wgrad(const void* dy_dev, const void* x_dev, int32_t M, int32_t N, int32_t K, float* d_weight_dev,
                       float* d_bias_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    MOLLY_CHECK(dy_dev && x_dev && d_weight_dev && workspace_dev, MOLLY_ERR_INVALID, "molly_linear_wgrad: NULL pointer");
    if (wgrad_through_transposes() && d_bias_dev != nullptr)
        return project_bwd_launch(dy_dev, x_dev, M, N, K, d_weight_dev, d_bias_dev, workspace_dev, workspace_bytes,
                                  static_cast<cudaStream_t>(stream));
    return linear_wgrad_launch(dy_dev, x_dev, M, N, K, d_weight_dev, d_bias_dev, static_cast<cudaStream_t>(stream));
}

int molly_gather_rows(void* d_hidden_dev, int32_t hs_dtype, const int32_t* seq_table_dev, int32_t n_seq, int32_t k_tokens,
                      int32_t k_cap, int32_t B, int32_t T, int32_t D, void* dy_dev, int32_t zero_rows, void* stream) {
    MOLLY_CHECK(d_hidden_dev && seq_table_dev && dy_dev, MOLLY_ERR_INVALID, "molly_gather_rows: NULL pointer");
    return gather_grad_rows_launch(d_hidden_dev, hs_dtype, seq_table_dev, n_seq, k_tokens, k_cap, B, T, D, dy_dev, zero_rows,
                                   static_cast<cudaStream_t>(stream));
}

int molly_transpose_bf16(const void* in_dev, int32_t rows, int32_t cols, void* out_dev, void* stream) {
    MOLLY_CHECK(in_dev && out_dev, MOLLY_ERR_INVALID, "molly_transpose_bf16: NULL pointer");
    return transpose_bf16_launch(in_dev, rows, cols, out_dev, static_cast<cudaStream_t>(stream));
}

int molly_layernorm_bwd(const float* x_dev, const void* dy_dev, const float* gamma_dev, int32_t rows, int32_t h, float eps,
                        float* d_x_dev, int32_t accumulate, float* stats_dev, float* d_gamma_dev, float* d_beta_dev,
                        void* stream) {
    MOLLY_CHECK(x_dev && dy_dev && gamma_dev && d_x_dev && stats_dev, MOLLY_ERR_INVALID, "molly_layernorm_bwd: NULL pointer");
    return ln_bwd_launch(x_dev, dy_dev, gamma_dev, rows, h, eps, d_x_dev, accumulate, stats_dev, d_gamma_dev, d_beta_dev,
                         static_cast<cudaStream_t>(stream));
}

int molly_act_fwd_bwd(int32_t glu, const void* pre_dev, const void* d_act_dev, int64_t rows, int32_t f_out, void* act_dev,
                      void* d_pre_dev, void* stream) {
    MOLLY_CHECK(pre_dev && act_dev && (d_act_dev == nullptr || d_pre_dev != nullptr), MOLLY_ERR_INVALID,
                "molly_act_fwd_bwd: NULL pointer");
    return act_fwd_bwd_launch(glu, pre_dev, d_act_dev, rows, f_out, act_dev, d_pre_dev, static_cast<cudaStream_t>(stream));
}

int molly_cast_f32_bf16(const float* in_dev, int64_t n, void* out_dev, void* stream) {
    MOLLY_CHECK(in_dev && out_dev, MOLLY_ERR_INVALID, "molly_cast_f32_bf16: NULL pointer");
    return cast_f32_bf16_launch(in_dev, n, out_dev, static_cast<cudaStream_t>(stream));
}

int molly_scale_cols(void* x_dev, int32_t rows, int32_t ld, int32_t cols, float scale, void* stream) {
    MOLLY_CHECK(x_dev, MOLLY_ERR_INVALID, "molly_scale_cols: NULL pointer");
    return scale_cols_launch(x_dev, rows, ld, cols, scale, static_cast<cudaStream_t>(stream));
}

int molly_scatter_add_rows(const float* src_dev, const int32_t* index_dev, const float* scale_dev, int32_t rows, int32_t h,
                           float* table_dev, void* stream) {
    MOLLY_CHECK(src_dev && index_dev && scale_dev && table_dev, MOLLY_ERR_INVALID, "molly_scatter_add_rows: NULL pointer");
    return scatter_add_rows_launch(src_dev, index_dev, scale_dev, rows, h, table_dev, static_cast<cudaStream_t>(stream));
}

// ------------------------------------ single kernels ------------------------------------
int molly_gemm_bf16(const void* a_dev, int32_t lda, const void* w_dev, int32_t ldw, int32_t M, int32_t N, int32_t K,
                    int32_t epilogue, const float* bias_dev, const float* residual_dev, void* out_dev,
                    int32_t out_dtype, int32_t ldo, const int32_t* seq_table_dev, int32_t seq_k_tokens, int32_t B,
                    int32_t T, int32_t k_cap, int32_t* err_flag_dev, int32_t scale_cols, float scale,
                    const float* rope_cos_t_dev, const float* rope_sin_t_dev, int32_t rope_len, int32_t rope_cols,
                    int32_t rope_head_dim, void* stream) {
    MOLLY_CHECK(a_dev && w_dev && out_dev, MOLLY_ERR_INVALID, "molly_gemm_bf16: NULL pointer");
    auto s = static_cast<cudaStream_t>(stream);
    CUtensorMap ta, tb, tc;
    int rc = gemm_make_map_a(&ta, a_dev, lda, M, K);
    if (rc) return rc;
    if ((rc = gemm_make_map_b(&tb, w_dev, ldw, N, K, epilogue))) return rc;
    CUtensorMap tr;
    const CUtensorMap* trp = nullptr;                                  // residual read in place unless it lives elsewhere
    if (epilogue == EPI_BIAS_RESID) {
        MOLLY_CHECK(residual_dev != nullptr, MOLLY_ERR_INVALID, "molly_gemm_bf16: residual epilogue needs a residual");
        if (static_cast<const void*>(residual_dev) != out_dev) {       // (same pitch as the output)
            if ((rc = gemm_make_map_c(&tr, const_cast<float*>(residual_dev), DT_F32, ldo, M, N))) return rc;
            trp = &tr;
        }
    }
    if (epilogue != EPI_SCATTER)
        if ((rc = gemm_make_map_c(&tc, out_dev, out_dtype, ldo, M, epilogue == EPI_GLU ? N / 2 : N))) return rc;
    return gemm_launch(ta, tb, epilogue == EPI_SCATTER ? nullptr : &tc, M, N, K, epilogue, bias_dev, out_dev, out_dtype,
                       ldo, seq_table_dev, seq_k_tokens, B, T, k_cap, err_flag_dev, s, scale_cols, scale, rope_cos_t_dev,
                       rope_sin_t_dev, rope_len, rope_cols, rope_head_dim, trp);
}

int molly_layernorm(const float* x_dev, const float* w_dev, const float* b_dev, int32_t rows, int32_t h, float eps,
                    void* out_dev, int32_t out_dtype, void* stream) {
    MOLLY_CHECK(x_dev && w_dev && b_dev && out_dev, MOLLY_ERR_INVALID, "molly_layernorm: NULL pointer");
    return layernorm_launch(x_dev, w_dev, b_dev, rows, h, eps, out_dev, out_dtype, static_cast<cudaStream_t>(stream));
}

int molly_embed(const int64_t* ids_dev, int32_t n_seq, int32_t k_tokens, const molly_encoder_config* cfg,
                const void* word_emb_dev, const void* pos_emb_dev, float* x_dev, int32_t* kv_info_dev,
                uint8_t* key_mask_dev, int32_t* err_flag_dev, void* stream) {
    MOLLY_CHECK(ids_dev && cfg && word_emb_dev && x_dev && kv_info_dev && key_mask_dev, MOLLY_ERR_INVALID,
                "molly_embed: NULL pointer");
    EmbedArgs ea{cfg->hidden_size, cfg->vocab_size, cfg->pad_token_id, cfg->mask_token_id, cfg->position_type,
                 cfg->max_positions, cfg->token_dropout, cfg->emb_layer_norm_before ? 0 : 1};
    return embed_launch(ids_dev, n_seq, k_tokens, ea, word_emb_dev, pos_emb_dev, x_dev, kv_info_dev, key_mask_dev,
                        err_flag_dev, static_cast<cudaStream_t>(stream));
}

int molly_rotary(void* qkv_dev, int32_t rows, int32_t k_tokens, int32_t h, int32_t heads, const float* cos_dev,
                 const float* sin_dev, void* stream) {
    MOLLY_CHECK(qkv_dev && cos_dev && sin_dev, MOLLY_ERR_INVALID, "molly_rotary: NULL pointer");
    return rotary_launch(qkv_dev, rows, k_tokens, h, heads, cos_dev, sin_dev, static_cast<cudaStream_t>(stream));
}

int molly_attention(const void* qkv_dev, int32_t n_seq, int32_t k_tokens, int32_t h, int32_t heads,
                    const int32_t* kv_info_dev, const uint8_t* key_mask_dev, void* out_dev, void* stream) {
    MOLLY_CHECK(qkv_dev && kv_info_dev && key_mask_dev && out_dev, MOLLY_ERR_INVALID, "molly_attention: NULL pointer");
    AttnMaps tm;
    int rc = attention_make_map(&tm, qkv_dev, n_seq * k_tokens, h, heads);
    if (rc) return rc;
    return attention_launch(tm, n_seq, k_tokens, h, heads, kv_info_dev, key_mask_dev, out_dev,
                            static_cast<cudaStream_t>(stream));
}

int molly_placeholder_runs(const int64_t* input_ids_dev, int32_t B, int32_t T, const int64_t pad_token_ids[3],
                           const int32_t* n_slots_dev, int32_t max_runs, int32_t* run_start_dev, int32_t* run_kind_dev, int32_t* run_len_dev,
                           int32_t* n_runs_dev, int32_t* pos_j_dev, void* stream) {
    MOLLY_CHECK(input_ids_dev && pad_token_ids && run_start_dev && run_kind_dev && run_len_dev && n_runs_dev && pos_j_dev,
                MOLLY_ERR_INVALID, "molly_placeholder_runs: NULL pointer");
    return placeholder_runs_launch(input_ids_dev, B, T, pad_token_ids[0], pad_token_ids[1], pad_token_ids[2], n_slots_dev,
                                   max_runs, run_start_dev, run_kind_dev, run_len_dev, n_runs_dev, pos_j_dev,
                                   static_cast<cudaStream_t>(stream));
}

int molly_placeholder_reject(int32_t* pos_j_dev, const int32_t* run_start_dev, const int32_t* run_kind_dev,
                             const int32_t* run_len_dev, const int32_t* n_runs_dev, const int32_t* slot_expect_dev, int32_t B,
                             int32_t T, int32_t max_runs, int32_t cap_dna_rna, int32_t cap_protein, void* stream) {
    MOLLY_CHECK(pos_j_dev && run_start_dev && run_kind_dev && run_len_dev && n_runs_dev && slot_expect_dev, MOLLY_ERR_INVALID,
                "molly_placeholder_reject: NULL pointer");
    return placeholder_reject_launch(pos_j_dev, run_start_dev, run_kind_dev, run_len_dev, n_runs_dev, slot_expect_dev, B, T,
                                     max_runs, cap_dna_rna, cap_protein, static_cast<cudaStream_t>(stream));
}

int molly_embed_tokens_skip(const int64_t* input_ids_dev, const int32_t* pos_j_dev, const int64_t pad_token_ids[3],
                            int32_t cap_dna_rna, int32_t cap_protein, const void* table_dev, int32_t dtype,
                            int32_t vocab, int32_t D, void* out_dev, int32_t B, int32_t T, int32_t* err_flag_dev,
                            void* stream) {
    MOLLY_CHECK(input_ids_dev && pos_j_dev && pad_token_ids && table_dev && out_dev, MOLLY_ERR_INVALID,
                "molly_embed_tokens_skip: NULL pointer");
    return embed_tokens_skip_launch(input_ids_dev, pos_j_dev, pad_token_ids[0], pad_token_ids[1], cap_dna_rna, cap_protein,
                                    table_dev, dtype, vocab, D, out_dev, B, T, err_flag_dev,
                                    static_cast<cudaStream_t>(stream));
}

int molly_build_seq_table(const int32_t* b_idx_dev, const int32_t* slot_idx_dev, int32_t n, const int32_t* run_start_dev,
                          const int32_t* run_kind_dev, const int32_t* run_len_dev, const int32_t* n_runs_dev,
                          int32_t max_runs, int32_t expect_protein, int32_t k_need, int32_t* seq_table_dev,
                          int32_t* err_flag_dev, void* stream) {
    MOLLY_CHECK(b_idx_dev && slot_idx_dev && run_start_dev && run_kind_dev && run_len_dev && n_runs_dev && seq_table_dev,
                MOLLY_ERR_INVALID, "molly_build_seq_table: NULL pointer");
    return build_seq_table_launch(b_idx_dev, slot_idx_dev, n, run_start_dev, run_kind_dev, run_len_dev, n_runs_dev,
                                  max_runs, expect_protein, k_need, seq_table_dev, err_flag_dev,
                                  static_cast<cudaStream_t>(stream));
}

int molly_attention_lse(const void* qkv_dev, int32_t n_seq, int32_t k_tokens, int32_t h, int32_t heads,
                        const int32_t* kv_info_dev, const uint8_t* key_mask_dev, void* out_dev, float* lse2_dev,
                        void* stream) {
    MOLLY_CHECK(qkv_dev && kv_info_dev && key_mask_dev && out_dev && lse2_dev, MOLLY_ERR_INVALID,
                "molly_attention_lse: NULL pointer");
    AttnMaps tm;
    int rc = attention_make_map(&tm, qkv_dev, n_seq * k_tokens, h, heads);
    if (rc) return rc;
    return attention_launch(tm, n_seq, k_tokens, h, heads, kv_info_dev, key_mask_dev, out_dev,
                            static_cast<cudaStream_t>(stream), lse2_dev);
}

int molly_attention_bwd(const void* qkv_dev, const void* out_dev, const void* d_out_dev, const float* lse2_dev, int32_t n_seq,
                        int32_t k_tokens, int32_t h, int32_t heads, const int32_t* kv_info_dev, const uint8_t* key_mask_dev,
                        void* d_qkv_dev, float* delta_ws_dev, void* stream) {
    MOLLY_CHECK(qkv_dev && out_dev && d_out_dev && lse2_dev && kv_info_dev && key_mask_dev && d_qkv_dev && delta_ws_dev,
                MOLLY_ERR_INVALID, "molly_attention_bwd: NULL pointer");
    return attention_bwd_launch(qkv_dev, out_dev, d_out_dev, lse2_dev, n_seq, k_tokens, h, heads, kv_info_dev, key_mask_dev,
                                d_qkv_dev, delta_ws_dev, static_cast<cudaStream_t>(stream));
}

int molly_attention_debug(long long* timeline_dev) {
    attention_set_debug(timeline_dev);
    return MOLLY_OK;
}

int molly_merge_rows(const void* src_dev, const int32_t* seq_table_dev, int32_t n_seq, int32_t k_tokens, int32_t k_cap,
                     void* hidden_states_dev, int32_t dtype, int32_t B, int32_t T, int32_t D, int32_t* err_flag_dev,
                     void* stream) {
    MOLLY_CHECK(src_dev && seq_table_dev && hidden_states_dev, MOLLY_ERR_INVALID, "molly_merge_rows: NULL pointer");
    return merge_rows_launch(src_dev, seq_table_dev, n_seq, k_tokens, k_cap, hidden_states_dev, dtype, B, T, D,
                             err_flag_dev, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
