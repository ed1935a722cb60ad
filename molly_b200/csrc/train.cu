// Native orchestration of the encoder TRAINING step (`--train-bio`: src/utils/tools.py:313-331 unfreezes
// the encoders): one C call runs the forward of every layer and keeps what the backward needs in a caller-owned tape; one
// C call runs the backward of a range of layers and leaves every parameter gradient in a flat fp32 buffer whose layout the
// library defines (molly_encoder_grad_layout).  The host enqueues ~27 launches per layer from C++ (the first version did this
// from Python, one ctypes call per kernel); once a shape is steady state the Python side captures the two calls into CUDA graphs.
//
// Autograd of the HF modules the kernels replace (HF EsmLayer, HF:446-482):
//     nn.Linear   dgrad / wgrad: the main tcgen05 GEMM with MN-major operands (gemm_launch_mn: the weight and the activations
//                 are read as they lie in memory, the wgrad is split over K); bias: column sums
//     attention   attention_bwd_launch (dQ / dK,dV kernels) from the forward's row log-sum-exp
//     LayerNorm   ln_bwd_launch        GELU / gated SiLU   act_fwd_bwd_launch
//     rotary      the forward kernel with -sin (inverse rotation) and the q scale folded in
#include <stdlib.h>

#include "common.h"
#include "encoder.h"
#include "kernels.h"

using namespace molly;

namespace {

constexpr size_t kAlign = 1024;
size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct Dims {
    int n_seq, k, M, h, H, F, F1, L;
    bool glu, rope, ffn_bias;
};

Dims dims_of(const molly_encoder* e, int n_seq, int k) {
    const auto& c = e->cfg;
    Dims d;
    d.n_seq = n_seq; d.k = k; d.M = n_seq * k; d.h = c.hidden_size; d.H = c.num_heads; d.F = c.intermediate_size;
    d.F1 = e->ffn1_n; d.L = c.num_layers;
    d.glu = c.ffn_type == MOLLY_FFN_GLU;
    d.rope = c.position_type == MOLLY_POS_ROTARY;
    d.ffn_bias = e->b_ffn1[0] != nullptr;
    return d;
}

// ---- tape: what the forward keeps.  x_in[l] (fp32 layer inputs, l = 0..L; x_in[L] feeds emb_layer_norm_after) always;
// the per-layer activations for every layer (recompute == 0) or one slot the backward refills layer by layer.
struct Tape {
    size_t x_in, x_stride, kv_info, key_mask, mid, slot0, slot_stride, total;
    size_t ln1, qkv, attn, lse2, x_mid, ln2, pre, act;     // offsets inside a slot
    int slots;
};

Tape tape_layout(const Dims& d, int recompute) {
    const size_t M = d.M, h = d.h;
    Tape t;
    size_t o = 0;
    t.x_stride = align_up(M * h * 4);
    t.x_in = o;     o += t.x_stride * (d.L + 1);
    t.kv_info = o;  o += align_up(static_cast<size_t>(d.n_seq) * 2 * 4);
    t.key_mask = o; o += align_up(M);
    t.mid = o;      o += align_up(M * d.F * 2);              // forward scratch (recompute mode): act(pre), the A operand of FFN2
    size_t s = 0;
    t.ln1 = s;   s += align_up(M * h * 2);
    t.qkv = s;   s += align_up(M * 3 * h * 2);
    t.attn = s;  s += align_up(M * h * 2);
    t.lse2 = s;  s += align_up(static_cast<size_t>(d.n_seq) * d.H * d.k * 4);
    t.x_mid = s; s += align_up(M * h * 4);
    t.ln2 = s;   s += align_up(M * h * 2);
    t.pre = s;   s += align_up(M * d.F1 * 2);
    t.act = s;   s += recompute ? 0 : align_up(M * d.F * 2);   // act(pre): the forward produces it anyway; kept, the backward
    t.slot_stride = s;                                          // neither recomputes nor rewrites it (recompute mode: scratch)
    t.slots = recompute ? 1 : d.L;
    t.slot0 = o;    o += s * t.slots;
    t.total = o;
    return t;
}

// ---- backward scratch
struct Scratch {
    size_t d_x, dy, d_act, act, d_pre, d_ln, d_attn, d_qkv, delta, stats, w_t, total;
};

Scratch scratch_layout(const Dims& d) {
    const size_t M = d.M, h = d.h;
    Scratch s;
    size_t o = 0;
    s.d_x = o;    o += align_up(M * h * 4);
    s.dy = o;     o += align_up(M * h * 2);
    s.d_act = o;  o += align_up(M * d.F * 2);
    s.act = o;    o += align_up(M * d.F * 2);
    s.d_pre = o;  o += align_up(M * d.F1 * 2);
    s.d_ln = o;   o += align_up(M * h * 2);
    s.d_attn = o; o += align_up(M * h * 2);
    s.d_qkv = o;  o += align_up(M * 3 * h * 2);
    s.delta = o;  o += align_up(static_cast<size_t>(d.n_seq) * d.H * d.k * 4);
    s.stats = o;  o += align_up(M * 2 * 4);
    const size_t wmax = static_cast<size_t>(d.F1 > 3 * d.h ? d.F1 : 3 * d.h) * h;
    s.w_t = o;    o += align_up(wmax * 2);
    s.total = o;
    return s;
}

// ---- gradient layout (floats).  Layer l's group starts at l * group; inside it the vectors come first (one memset), then
// the matrices, all in the PACKED layout of the weights (q,k,v concatenated; GLU rows interleaved).
struct GradLayout {
    long long off[MOLLY_GRAD_SLOTS];      // inside a layer group; -1 = the encoder has no such parameter
    long long vec_floats, group;          // floats of the vector block / of the whole layer group
    long long tail[MOLLY_GRAD_TAIL_SLOTS];
    long long total;
};

GradLayout grad_layout(const molly_encoder* e) {
    const Dims d = dims_of(e, 1, 1);
    const long long h = d.h, F = d.F, F1 = d.F1;
    GradLayout g;
    long long o = 0;
    auto put = [&](int slot, long long n, bool present) { g.off[slot] = present ? o : -1; if (present) o += n; };
    put(MOLLY_GRAD_LN2_W, h, true);        put(MOLLY_GRAD_LN2_B, h, true);
    put(MOLLY_GRAD_B_FFN2, h, d.ffn_bias); put(MOLLY_GRAD_B_FFN1, F1, d.ffn_bias);
    put(MOLLY_GRAD_LN1_W, h, true);        put(MOLLY_GRAD_LN1_B, h, true);
    put(MOLLY_GRAD_B_O, h, true);          put(MOLLY_GRAD_B_QKV, 3 * h, true);
    g.vec_floats = o;
    put(MOLLY_GRAD_W_FFN2, h * F, true);   put(MOLLY_GRAD_W_FFN1, F1 * h, true);
    put(MOLLY_GRAD_W_O, h * h, true);      put(MOLLY_GRAD_W_QKV, 3 * h * h, true);
    g.group = o;
    o = g.group * d.L;
    g.tail[MOLLY_GRAD_TAIL_FINAL_LN_W] = o; o += h;
    g.tail[MOLLY_GRAD_TAIL_FINAL_LN_B] = o; o += h;
    g.tail[MOLLY_GRAD_TAIL_WORD_EMB] = o;   o += static_cast<long long>(e->cfg.vocab_size) * h;
    g.tail[MOLLY_GRAD_TAIL_POS_EMB] = e->w.pos_emb_dev ? o : -1;
    if (e->w.pos_emb_dev) o += static_cast<long long>(e->cfg.max_positions) * h;
    g.total = o;
    return g;
}

int check_call(const molly_encoder* e, int n_seq, int k, const void* tape, size_t tape_bytes, int recompute) {
    MOLLY_CHECK(e != nullptr, MOLLY_ERR_INVALID, "train: NULL encoder");
    MOLLY_CHECK(!e->cfg.emb_layer_norm_before, MOLLY_ERR_UNSUPPORTED,
                "train: emb_layer_norm_before encoders are not covered by the training path");
    MOLLY_CHECK(n_seq > 0 && k > 0, MOLLY_ERR_INVALID, "train: n_seq=%d k_tokens=%d", n_seq, k);
    MOLLY_CHECK(tape != nullptr && (reinterpret_cast<uintptr_t>(tape) & (kAlign - 1)) == 0, MOLLY_ERR_INVALID,
                "train: the tape must be a 1024-B aligned device buffer");
    const size_t need = tape_layout(dims_of(e, n_seq, k), recompute).total;
    MOLLY_CHECK(tape_bytes >= need, MOLLY_ERR_WORKSPACE, "train: tape %zu B < required %zu B", tape_bytes, need);
    if (e->cfg.position_type == MOLLY_POS_ROTARY)
        MOLLY_CHECK(e->w.rope_len >= k && e->w.rope_cos_dev && e->w.rope_sin_dev, MOLLY_ERR_INVALID,
                    "rotary tables cover %d positions < k_tokens %d", e->w.rope_len, k);
    return MOLLY_OK;
}

// out[M, N] = a[M, K] w[N, K]^T (+ bias), bf16 out; maps built per call (a few microseconds of host time)
int gemm_plain(const void* a, const void* w, const CUtensorMap* tm_w, int M, int N, int K, int epi, const float* bias, void* out,
               int n_out, cudaStream_t s, const molly_encoder* e = nullptr, int seq_k = 0) {
    CUtensorMap ta, tb, tc;
    int rc = gemm_make_map_a(&ta, a, K, M, K);
    if (rc) return rc;
    if (tm_w == nullptr) {
        if ((rc = gemm_make_map_b(&tb, w, K, N, K, epi))) return rc;
        tm_w = &tb;
    }
    if ((rc = gemm_make_map_c(&tc, out, DT_BF16, n_out, M, n_out))) return rc;
    if (epi == EPI_BIAS_ROPE) {
        const int h = e->cfg.hidden_size, d = h / e->cfg.num_heads;
        return gemm_launch(ta, *tm_w, &tc, M, N, K, epi, bias, out, DT_BF16, n_out, nullptr, seq_k, 0, 0, 0, nullptr, s, h,
                           e->q_scale, e->w.rope_cos_t_dev, e->w.rope_sin_t_dev, e->w.rope_len, 2 * h, d);
    }
    return gemm_launch(ta, *tm_w, &tc, M, N, K, epi, bias, out, DT_BF16, n_out, nullptr, 0, 0, 0, 0, nullptr, s);
}

// x_out[M, h] (fp32) = x_res + a[M, K] w[h, K]^T + bias; the epilogue reads x_res through its own tensor map, so the layer
// input the tape keeps is never copied
int gemm_residual(const void* a, const CUtensorMap& tm_w, int M, int h, int K, const float* bias, const float* x_res,
                  float* x_out, cudaStream_t s) {
    CUtensorMap ta, tc, tr;
    int rc = gemm_make_map_a(&ta, a, K, M, K);
    if (rc) return rc;
    if ((rc = gemm_make_map_c(&tc, x_out, DT_F32, h, M, h))) return rc;
    if ((rc = gemm_make_map_c(&tr, const_cast<float*>(x_res), DT_F32, h, M, h))) return rc;
    return gemm_launch(ta, tm_w, &tc, M, h, K, EPI_BIAS_RESID, bias, x_out, DT_F32, h, nullptr, 0, 0, 0, 0, nullptr, s, 0, 1.0f,
                       nullptr, nullptr, 0, 0, 0, &tr);
}

// Layer l from x_in up to the FFN pre-activation, everything the backward needs written into `slot`.
int layer_forward_keep(const molly_encoder* e, const Dims& d, int l, const float* x_in, uint8_t* slot, const Tape& t,
                       const int32_t* kv_info, const uint8_t* key_mask, cudaStream_t s) {
    const auto& c = e->cfg;
    const int M = d.M, h = d.h, hd = h / d.H;
    void* ln1 = slot + t.ln1;
    void* qkv = slot + t.qkv;
    void* attn = slot + t.attn;
    float* lse2 = reinterpret_cast<float*>(slot + t.lse2);
    float* x_mid = reinterpret_cast<float*>(slot + t.x_mid);
    void* ln2 = slot + t.ln2;
    void* pre = slot + t.pre;
    int rc;
    if ((rc = layernorm_launch(x_in, e->ln1_w[l], e->ln1_b[l], M, h, c.layer_norm_eps, ln1, DT_BF16, s))) return rc;
    const bool rope_fused = d.rope && hd <= 64 && e->w.rope_cos_t_dev != nullptr && e->w.rope_sin_t_dev != nullptr;
    if (rope_fused) {
        if ((rc = gemm_plain(ln1, nullptr, &e->tm_wqkv[l], M, 3 * h, h, EPI_BIAS_ROPE, e->b_qkv[l], qkv, 3 * h, s, e, d.k)))
            return rc;
    } else {
        CUtensorMap ta, tc;
        if ((rc = gemm_make_map_a(&ta, ln1, h, M, h))) return rc;
        if ((rc = gemm_make_map_c(&tc, qkv, DT_BF16, 3 * h, M, 3 * h))) return rc;
        if ((rc = gemm_launch(ta, e->tm_wqkv[l], &tc, M, 3 * h, h, EPI_BIAS, e->b_qkv[l], qkv, DT_BF16, 3 * h, nullptr, 0, 0, 0,
                              0, nullptr, s, h, e->q_scale)))
            return rc;
        if (d.rope && (rc = rotary_launch(qkv, M, d.k, h, d.H, e->w.rope_cos_dev, e->w.rope_sin_dev, s))) return rc;
    }
    AttnMaps am;
    if ((rc = attention_make_map(&am, qkv, M, h, d.H))) return rc;
    if ((rc = attention_launch(am, d.n_seq, d.k, h, d.H, kv_info, key_mask, attn, s, lse2))) return rc;
    if ((rc = gemm_residual(attn, e->tm_wo[l], M, h, h, e->b_o[l], x_in, x_mid, s))) return rc;
    if ((rc = layernorm_launch(x_mid, e->ln2_w[l], e->ln2_b[l], M, h, c.layer_norm_eps, ln2, DT_BF16, s))) return rc;
    // the backward needs the pre-activation: plain bias epilogue, activation apart
    return gemm_plain(ln2, nullptr, &e->tm_w1[l], M, d.F1, h, EPI_BIAS, e->b_ffn1[l], pre, d.F1, s);
}

int forward(const molly_encoder* e, const int64_t* ids, int n_seq, int k, void* out, uint8_t* tape, int recompute,
            int32_t* err_flag, cudaStream_t s) {
    const auto& c = e->cfg;
    const Dims d = dims_of(e, n_seq, k);
    const Tape t = tape_layout(d, recompute);
    const int M = d.M, h = d.h;
    auto x_at = [&](int l) { return reinterpret_cast<float*>(tape + t.x_in + t.x_stride * l); };
    int32_t* kv_info = reinterpret_cast<int32_t*>(tape + t.kv_info);
    uint8_t* key_mask = tape + t.key_mask;
    void* mid = tape + t.mid;
    EmbedArgs ea{h, c.vocab_size, c.pad_token_id, c.mask_token_id, c.position_type, c.max_positions, c.token_dropout, 1};
    int rc = embed_launch(ids, n_seq, k, ea, e->w.word_emb_dev, e->w.pos_emb_dev, x_at(0), kv_info, key_mask, err_flag, s);
    if (rc) return rc;
    for (int l = 0; l < d.L; ++l) {
        uint8_t* slot = tape + t.slot0 + (recompute ? 0 : t.slot_stride * l);
        if ((rc = layer_forward_keep(e, d, l, x_at(l), slot, t, kv_info, key_mask, s))) return rc;
        // finish the layer: FFN activation + FFN2 on top of x_mid
        void* act = recompute ? mid : static_cast<void*>(slot + t.act);
        if ((rc = act_fwd_bwd_launch(d.glu, slot + t.pre, nullptr, M, d.F, act, nullptr, s))) return rc;
        if ((rc = gemm_residual(act, e->tm_w2[l], M, h, d.F, e->b_ffn2[l], reinterpret_cast<float*>(slot + t.x_mid),
                                x_at(l + 1), s)))
            return rc;
    }
    // emb_layer_norm_after (HF:511-512) -> hidden_states[-1]
    return layernorm_launch(x_at(d.L), e->w.final_ln_w_dev, e->w.final_ln_b_dev, M, h, c.layer_norm_eps, out, DT_BF16, s);
}

// dgrad of nn.Linear: d_in[M, K] = d_out[M, N] W[N, K], W read as it is stored ([N, K] row-major = MN-major second operand).
// MOLLY_DGRAD_TRANSPOSE=1: the first version (W^T materialised in scratch, then the K-major GEMM).
int dgrad(const void* d_out, const void* w, int M, int N, int K, void* w_t, void* d_in, cudaStream_t s) {
    static const bool legacy = [] { const char* e = getenv("MOLLY_DGRAD_TRANSPOSE"); return e != nullptr && e[0] == '1'; }();
    if (!legacy) return gemm_launch_mn(GEMM_OPND_K_MN, d_out, N, w, K, M, K, N, d_in, DT_BF16, K, s);
    int rc = transpose_bf16_launch(w, N, K, w_t, s);
    if (rc) return rc;
    return gemm_plain(d_out, w_t, nullptr, M, K, N, EPI_BIAS, nullptr, d_in, K, s);
}

int backward_layer(const molly_encoder* e, const Dims& d, int l, const float* x_in, const uint8_t* slot, const Tape& t,
                   const int32_t* kv_info, const uint8_t* key_mask, uint8_t* ws, const Scratch& sc, float* g, const GradLayout& gl,
                   bool act_kept, cudaStream_t s) {
    const auto& c = e->cfg;
    const int M = d.M, h = d.h, F = d.F, F1 = d.F1;
    float* d_x = reinterpret_cast<float*>(ws + sc.d_x);
    void* dy = ws + sc.dy;
    void* d_act = ws + sc.d_act;
    void* act = ws + sc.act;
    void* d_pre = ws + sc.d_pre;
    void* d_ln = ws + sc.d_ln;
    void* d_attn = ws + sc.d_attn;
    void* d_qkv = ws + sc.d_qkv;
    float* delta = reinterpret_cast<float*>(ws + sc.delta);
    float* stats = reinterpret_cast<float*>(ws + sc.stats);
    void* w_t = ws + sc.w_t;
    auto G = [&](int slot_id) { return gl.off[slot_id] < 0 ? nullptr : g + gl.off[slot_id]; };
    int rc;
    // On entry `dy` holds the bf16 copy of d_x = d(x_out) (written by the LayerNorm backward above this layer), this layer's
    // whole gradient group is zeroed (ONE memset: LayerNorm / bias gradients accumulate with atomics, split-K weight gradients
    // with TMA reduce-adds) and its b_ffn2 gradient (column sums of dy) is already in place.
    // ---- feed-forward block: x_out = x_mid + W2 act(W1 LN2(x_mid) + b1) + b2
    if ((rc = dgrad(dy, e->w_ffn2[l], M, h, F, w_t, d_act, s))) return rc;
    // d_pre = d_act * act'(pre); act itself comes from the tape when the forward kept it, else it is rebuilt in the same pass
    const void* act_in = act_kept ? static_cast<const void*>(slot + t.act) : act;
    if ((rc = act_fwd_bwd_launch(d.glu, slot + t.pre, d_act, M, F, act_kept ? nullptr : act, d_pre, s))) return rc;
    if ((rc = linear_wgrad_launch(dy, act_in, M, h, F, G(MOLLY_GRAD_W_FFN2), nullptr, s, true))) return rc;
    // (the GELU backward is bound by instruction issue: with the b_ffn1 column sums fused in it took 92 us against 45 + 17 us
    //  for the plain kernel and a separate column-sum pass over d_pre, so the bias gradient stays with the wgrad)
    if ((rc = linear_wgrad_launch(d_pre, slot + t.ln2, M, F1, h, G(MOLLY_GRAD_W_FFN1), G(MOLLY_GRAD_B_FFN1), s, true))) return rc;
    if ((rc = dgrad(d_pre, e->w_ffn1[l], M, F1, h, w_t, d_ln, s))) return rc;
    // d_x becomes d(x_mid); dy its bf16 copy; b_o gradient = column sums of dy
    if ((rc = ln_bwd_launch(reinterpret_cast<const float*>(slot + t.x_mid), d_ln, e->ln2_w[l], M, h, c.layer_norm_eps, d_x, 1,
                            stats, G(MOLLY_GRAD_LN2_W), G(MOLLY_GRAD_LN2_B), s, dy, G(MOLLY_GRAD_B_O))))
        return rc;
    // ---- attention block: x_mid = x_in + Wo Attn(LN1(x_in)) + bo
    if ((rc = dgrad(dy, e->w_o[l], M, h, h, w_t, d_attn, s))) return rc;
    if ((rc = linear_wgrad_launch(dy, slot + t.attn, M, h, h, G(MOLLY_GRAD_W_O), nullptr, s, true))) return rc;
    if ((rc = attention_bwd_launch(slot + t.qkv, slot + t.attn, d_attn, reinterpret_cast<const float*>(slot + t.lse2), d.n_seq,
                                   d.k, h, d.H, kv_info, key_mask, d_qkv, delta, s)))
        return rc;
    // inverse rotation of d(q'), d(k') and q = (W_q x + b_q) * d^-1/2 (HF:341)
    if (d.rope) {
        if ((rc = rotary_launch(d_qkv, M, d.k, h, d.H, e->w.rope_cos_dev, e->w.rope_sin_dev, s, -1.0f, e->q_scale))) return rc;
    } else if ((rc = scale_cols_launch(d_qkv, M, 3 * h, h, e->q_scale, s))) {
        return rc;
    }
    if ((rc = linear_wgrad_launch(d_qkv, slot + t.ln1, M, 3 * h, h, G(MOLLY_GRAD_W_QKV), G(MOLLY_GRAD_B_QKV), s, true))) return rc;
    if ((rc = dgrad(d_qkv, e->w_qkv[l], M, 3 * h, h, w_t, d_ln, s))) return rc;
    // d_x becomes d(x_in) = d(x_out of layer l-1): prepare that layer's vector block and its b_ffn2 gradient
    float* g_below = l > 0 ? g - gl.group : nullptr;
    if (g_below != nullptr) MOLLY_CUDA(cudaMemsetAsync(g_below, 0, sizeof(float) * gl.group, s));
    float* b2_below = (g_below != nullptr && gl.off[MOLLY_GRAD_B_FFN2] >= 0) ? g_below + gl.off[MOLLY_GRAD_B_FFN2] : nullptr;
    return ln_bwd_launch(x_in, d_ln, e->ln1_w[l], M, h, c.layer_norm_eps, d_x, 1, stats, G(MOLLY_GRAD_LN1_W), G(MOLLY_GRAD_LN1_B),
                         s, l > 0 ? dy : nullptr, b2_below);
}

}  // namespace

extern "C" {

int molly_encoder_train_sizes(const molly_encoder_t* enc, int32_t n_seq, int32_t k_tokens, int32_t recompute,
                              size_t* tape_bytes, size_t* workspace_bytes, int64_t* grad_floats) {
    MOLLY_CHECK(enc && n_seq > 0 && k_tokens > 0, MOLLY_ERR_INVALID, "molly_encoder_train_sizes: bad argument");
    const Dims d = dims_of(enc, n_seq, k_tokens);
    if (tape_bytes) *tape_bytes = tape_layout(d, recompute).total;
    if (workspace_bytes) *workspace_bytes = scratch_layout(d).total;
    if (grad_floats) *grad_floats = grad_layout(enc).total;
    return MOLLY_OK;
}

int molly_encoder_grad_layout(const molly_encoder_t* enc, int64_t* layer_offsets, int64_t* layer_group_floats,
                              int64_t* tail_offsets, int64_t* total_floats) {
    MOLLY_CHECK(enc && layer_offsets && layer_group_floats && tail_offsets && total_floats, MOLLY_ERR_INVALID,
                "molly_encoder_grad_layout: NULL argument");
    const GradLayout g = grad_layout(enc);
    for (int i = 0; i < MOLLY_GRAD_SLOTS; ++i) layer_offsets[i] = g.off[i];
    for (int i = 0; i < MOLLY_GRAD_TAIL_SLOTS; ++i) tail_offsets[i] = g.tail[i];
    *layer_group_floats = g.group;
    *total_floats = g.total;
    return MOLLY_OK;
}

int molly_encode_train_fwd(molly_encoder_t* enc, const int64_t* ids_dev, int32_t n_seq, int32_t k_tokens, void* out_dev,
                           void* tape_dev, size_t tape_bytes, int32_t recompute, int32_t* err_flag_dev, void* stream) {
    int rc = check_call(enc, n_seq, k_tokens, tape_dev, tape_bytes, recompute);
    if (rc) return rc;
    MOLLY_CHECK(ids_dev && out_dev, MOLLY_ERR_INVALID, "molly_encode_train_fwd: NULL pointer");
    set_gemm_family(PF_GEMM_OTHER);
    return forward(enc, ids_dev, n_seq, k_tokens, out_dev, static_cast<uint8_t*>(tape_dev), recompute, err_flag_dev,
                   static_cast<cudaStream_t>(stream));
}

int molly_encode_train_bwd(molly_encoder_t* enc, int32_t n_seq, int32_t k_tokens, void* tape_dev, size_t tape_bytes,
                           int32_t recompute, const void* d_out_dev, float* grads_dev, void* workspace_dev,
                           size_t workspace_bytes, int32_t layer_begin, int32_t layer_end, void* stream) {
    int rc = check_call(enc, n_seq, k_tokens, tape_dev, tape_bytes, recompute);
    if (rc) return rc;
    const Dims d = dims_of(enc, n_seq, k_tokens);
    const Tape t = tape_layout(d, recompute);
    const Scratch sc = scratch_layout(d);
    const GradLayout gl = grad_layout(enc);
    MOLLY_CHECK(grads_dev && workspace_dev && (reinterpret_cast<uintptr_t>(workspace_dev) & (kAlign - 1)) == 0 &&
                    (reinterpret_cast<uintptr_t>(grads_dev) & 15) == 0, MOLLY_ERR_INVALID,
                "molly_encode_train_bwd: NULL / misaligned buffer");
    MOLLY_CHECK(workspace_bytes >= sc.total, MOLLY_ERR_WORKSPACE, "molly_encode_train_bwd: workspace %zu B < required %zu B",
                workspace_bytes, sc.total);
    MOLLY_CHECK(layer_begin <= d.L && layer_end >= 0 && layer_end <= layer_begin, MOLLY_ERR_INVALID,
                "molly_encode_train_bwd: layers %d .. %d outside %d .. 0", layer_begin, layer_end, d.L);
    auto s = static_cast<cudaStream_t>(stream);
    auto* tape = static_cast<uint8_t*>(tape_dev);
    auto* ws = static_cast<uint8_t*>(workspace_dev);
    auto x_at = [&](int l) { return reinterpret_cast<float*>(tape + t.x_in + t.x_stride * l); };
    const int32_t* kv_info = reinterpret_cast<const int32_t*>(tape + t.kv_info);
    const uint8_t* key_mask = tape + t.key_mask;
    float* d_x = reinterpret_cast<float*>(ws + sc.d_x);
    set_gemm_family(PF_GEMM_OTHER);
    for (int l = layer_begin; l >= layer_end; --l) {
        if (l == d.L) {                            // emb_layer_norm_after: d_x = LN-backward(d_out)
            MOLLY_CHECK(d_out_dev != nullptr, MOLLY_ERR_INVALID, "molly_encode_train_bwd: d_out is needed for layer %d", d.L);
            float* gw = grads_dev + gl.tail[MOLLY_GRAD_TAIL_FINAL_LN_W];
            float* g_top = grads_dev + gl.group * (d.L - 1);          // layer L-1: zero its vector block, fill its b_ffn2 gradient
            MOLLY_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * 2 * d.h, s));
            MOLLY_CUDA(cudaMemsetAsync(g_top, 0, sizeof(float) * gl.group, s));
            if ((rc = ln_bwd_launch(x_at(d.L), d_out_dev, enc->w.final_ln_w_dev, d.M, d.h, enc->cfg.layer_norm_eps, d_x, 0,
                                    reinterpret_cast<float*>(ws + sc.stats), gw, gw + d.h, s, ws + sc.dy,
                                    gl.off[MOLLY_GRAD_B_FFN2] >= 0 ? g_top + gl.off[MOLLY_GRAD_B_FFN2] : nullptr)))
                return rc;
            continue;
        }
        uint8_t* slot = tape + t.slot0 + (recompute ? 0 : t.slot_stride * l);
        if (recompute && (rc = layer_forward_keep(enc, d, l, x_at(l), slot, t, kv_info, key_mask, s))) return rc;
        if ((rc = backward_layer(enc, d, l, x_at(l), slot, t, kv_info, key_mask, ws, sc, grads_dev + gl.group * l, gl, !recompute,
                                 s)))
            return rc;
    }
    return MOLLY_OK;
}

}  // extern "C"
