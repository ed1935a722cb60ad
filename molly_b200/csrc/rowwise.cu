// HBM-bound row-wise kernels of the encoder: embedding gather (+ token-dropout, learned positions, pad masking),
// LayerNorm, rotary, pooled read-outs.  All use 16-byte vector accesses, one warp per row, warp-shuffle reductions.
#include <math_constants.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// Embedding gather.  HF:189-236 (EsmEmbeddings.forward) + HF:971-984 (position ids) + omics_one.py:70 (mask = ids != 1)
//   x = word_emb[id]; token_dropout: zero <mask> rows, x = x * 0.88 / (1 - n_mask / n_nonpad);
//   absolute: x += pos_emb[cumsum(mask) * mask + pad_idx];  x *= mask
// grid = (n_seq, splits); every block re-derives the per-sequence statistics (K int64 reads, L2 resident) and then
// gathers its slice of the K token rows, one warp per row.
// ------------------------------------------------------------------------------------------------
constexpr int EMB_THREADS = 256;

__global__ void __launch_bounds__(EMB_THREADS)
embed_kernel(const int64_t* __restrict__ ids, int k_tokens, EmbedArgs a, const __nv_bfloat16* __restrict__ word_emb,
             const __nv_bfloat16* __restrict__ pos_emb, float* __restrict__ x, int32_t* __restrict__ kv_len,
             uint8_t* __restrict__ key_mask, int32_t* err_flag) {
    const int apply_mask = a.apply_mask;
    extern __shared__ int32_t s_pos[];                 // [k_tokens] position ids (absolute) -- or unused
    __shared__ int s_warp_cnt[EMB_THREADS / 32];
    __shared__ int s_red[3][EMB_THREADS / 32];
    __shared__ int s_base;
    const int n = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t* seq = ids + static_cast<size_t>(n) * k_tokens;

    // pass 1: statistics + (absolute) running non-pad count
    int n_nonpad = 0, n_mask = 0, last = 0;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int t0 = 0; t0 < k_tokens; t0 += EMB_THREADS) {
        const int t = t0 + threadIdx.x;
        int64_t id = (t < k_tokens) ? seq[t] : 1;
        const bool nonpad = (t < k_tokens) && (id != 1);          // reference hard-codes pad id 1 (omics_one.py:70)
        if (t < k_tokens) {
            if (id < 0 || id >= a.vocab) { if (err_flag) atomicOr(err_flag, 1); }
            n_nonpad += nonpad;
            n_mask += (id == a.mask_id);
            if (nonpad) last = t + 1;
            if (blockIdx.y == 0) key_mask[static_cast<size_t>(n) * k_tokens + t] = nonpad ? 1 : 0;
        }
        if (a.position_type == 1) {                                // cumsum(mask) * mask + pad_idx
            const unsigned bal = __ballot_sync(0xffffffffu, nonpad);
            const int incl = __popc(bal & (0xffffffffu >> (31 - lane)));
            if (lane == 31) s_warp_cnt[warp] = incl;
            __syncthreads();
            int before = s_base;
            for (int w = 0; w < warp; ++w) before += s_warp_cnt[w];
            if (t < k_tokens) s_pos[t] = nonpad ? (before + incl + a.pad_id) : a.pad_id;
            __syncthreads();
            if (threadIdx.x == EMB_THREADS - 1) s_base = before + incl;
            __syncthreads();
        }
    }
    n_nonpad = warp_sum_i(n_nonpad);
    n_mask = warp_sum_i(n_mask);
    last = warp_max_i(last);
    if (lane == 0) { s_red[0][warp] = n_nonpad; s_red[1][warp] = n_mask; s_red[2][warp] = last; }
    __syncthreads();
    n_nonpad = 0; n_mask = 0; last = 0;
    for (int w = 0; w < EMB_THREADS / 32; ++w) {
        n_nonpad += s_red[0][w]; n_mask += s_red[1][w]; last = max(last, s_red[2][w]);
    }
    if (blockIdx.y == 0 && threadIdx.x == 0) {      // kv_info[n] = (last non-pad index + 1, number of non-pad ids)
        kv_len[2 * n] = last;
        kv_len[2 * n + 1] = n_nonpad;
    }

    // HF:213-222: (x * (1 - 0.15*0.8)) / (1 - n_mask / src_len)
    const float keep = 0.88f;                                      // python: 1 - 0.15 * 0.8 -> float32(0.88)
    const float denom = 1.0f - static_cast<float>(n_mask) / static_cast<float>(n_nonpad);

    // pass 2: gather this block's slice of rows, one warp per row
    const int per = (k_tokens + gridDim.y - 1) / gridDim.y;
    const int t_begin = blockIdx.y * per, t_end = min(k_tokens, t_begin + per);
    const int nvec = a.hidden / 8;                                 // 8 bf16 = 16 B per vector
    for (int t = t_begin + warp; t < t_end; t += EMB_THREADS / 32) {
        int64_t id = seq[t];
        const bool nonpad = id != 1;
        if (id < 0 || id >= a.vocab) id = a.pad_id;
        const bool zero_row = (apply_mask && !nonpad);
        const bool drop = a.token_dropout && (id == a.mask_id);
        int pos = a.pad_id;
        if (a.position_type == 1) {
            pos = s_pos[t];
            if (pos >= a.max_positions) { if (err_flag && lane == 0) atomicOr(err_flag, 4); pos = a.pad_id; }
        }
        const uint4* src = reinterpret_cast<const uint4*>(word_emb + static_cast<size_t>(id) * a.hidden);
        const uint4* psrc = reinterpret_cast<const uint4*>(pos_emb + static_cast<size_t>(pos) * a.hidden);
        float4* dst = reinterpret_cast<float4*>(x + (static_cast<size_t>(n) * k_tokens + t) * a.hidden);
        for (int v = lane; v < nvec; v += 32) {
            float f[8];
            if (zero_row) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = 0.f;
            } else {
                const uint4 u = __ldg(src + v);
                const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 p = __bfloat1622float2(b2[i]);
                    f[2 * i] = p.x; f[2 * i + 1] = p.y;
                }
                if (a.token_dropout) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = drop ? 0.f : (f[i] * keep) / denom;
                }
                if (a.position_type == 1) {
                    const uint4 pu = __ldg(psrc + v);
                    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pu);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 p = __bfloat1622float2(p2[i]);
                        f[2 * i] += p.x; f[2 * i + 1] += p.y;
                    }
                }
            }
            dst[2 * v] = make_float4(f[0], f[1], f[2], f[3]);
            dst[2 * v + 1] = make_float4(f[4], f[5], f[6], f[7]);
        }
    }
}

__global__ void mask_rows_kernel(float* __restrict__ x, const uint8_t* __restrict__ key_mask, int rows, int h) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows || key_mask[row]) return;
    float4* r = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * h);
    for (int v = threadIdx.x & 31; v < h / 4; v += 32) r[v] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (HF:394, 479, 511-512): fp32 residual stream in, bf16 (GEMM A operand) or fp32 out.
// One warp per row; the row is held in registers (<= 20 float4 per lane, h <= 2560); mean, then sum of squared
// deviations (two-pass, like torch) by warp shuffles; one HBM read + one write per element.
// ------------------------------------------------------------------------------------------------
constexpr int LN_THREADS = 128;                 // 4 rows per CTA: small CTAs keep many loads in flight per SM

template <typename OutT, int VPL>               // VPL = float4 vectors per lane = ceil(h / 128)
__global__ void __launch_bounds__(LN_THREADS)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int rows, int h,
                 float eps, OutT* __restrict__ out) {
    const int row = blockIdx.x * (LN_THREADS / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int nvec = h >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * h);
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        v[i] = (idx < nvec) ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / static_cast<float>(h);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        if (lane + 32 * i < nvec) {
            const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
            ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(h) + eps);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            const float4 ww = __ldg(w4 + idx), bb = __ldg(b4 + idx);
            const float y0 = (v[i].x - mean) * rstd * ww.x + bb.x;
            const float y1 = (v[i].y - mean) * rstd * ww.y + bb.y;
            const float y2 = (v[i].z - mean) * rstd * ww.z + bb.z;
            const float y3 = (v[i].w - mean) * rstd * ww.w + bb.w;
            if constexpr (sizeof(OutT) == 2) {
                uint2 u;
                u.x = pack_bf16x2(y0, y1);
                u.y = pack_bf16x2(y2, y3);
                reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * h)[idx] = u;
            } else {
                reinterpret_cast<float4*>(out + static_cast<size_t>(row) * h)[idx] = make_float4(y0, y1, y2, y3);
            }
        }
    }
}

template <typename OutT>
void launch_layernorm(const float* x, const float* w, const float* b, int rows, int h, float eps, OutT* out,
                      cudaStream_t stream) {
    const int grid = (rows + LN_THREADS / 32 - 1) / (LN_THREADS / 32);
    const int vpl = (h / 4 + 31) / 32;
#define MOLLY_LN_CASE(V) \
    if (vpl <= V) { layernorm_kernel<OutT, V><<<grid, LN_THREADS, 0, stream>>>(x, w, b, rows, h, eps, out); return; }
    MOLLY_LN_CASE(1) MOLLY_LN_CASE(2) MOLLY_LN_CASE(3) MOLLY_LN_CASE(4) MOLLY_LN_CASE(5) MOLLY_LN_CASE(6)
    MOLLY_LN_CASE(8) MOLLY_LN_CASE(10) MOLLY_LN_CASE(12) MOLLY_LN_CASE(16) MOLLY_LN_CASE(20)
#undef MOLLY_LN_CASE
}

// ------------------------------------------------------------------------------------------------
// Rotary (HF:45-54, 81-123): in place on the q and k thirds of the packed bf16 [rows, 3h] QKV buffer.
// angle = (row index inside the padded sequence) * inv_freq; half-split (NeoX) pairing (i, i + d/2);
// cos/sin come from fp32 tables [k_tokens, d/2].  Each thread rotates 8 pairs (two 16-B vectors).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rotary_kernel(__nv_bfloat16* __restrict__ qkv, long long total_items, int k_tokens, int h, int heads,
              const float* __restrict__ cos_t, const float* __restrict__ sin_t, float sin_sign, float q_scale) {
    const int d = h / heads, half = d >> 1, vec_per_head = half >> 3;
    const long long item = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (item >= total_items) return;
    const int per_row = 2 * heads * vec_per_head;
    const long long row = item / per_row;
    int rem = static_cast<int>(item - row * per_row);
    const int which = rem / (heads * vec_per_head);           // 0 = q, 1 = k
    rem -= which * heads * vec_per_head;
    const int head = rem / vec_per_head, vec = rem - head * vec_per_head;
    const int t = static_cast<int>(row % k_tokens);
    __nv_bfloat16* base = qkv + row * (3LL * h) + which * h + head * d + vec * 8;
    uint4 lo = *reinterpret_cast<uint4*>(base), hi = *reinterpret_cast<uint4*>(base + half);
    const float4* c4 = reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(t) * half + vec * 8);
    const float4* s4 = reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(t) * half + vec * 8);
    float c[8], s[8];
    *reinterpret_cast<float4*>(c) = __ldg(c4); *reinterpret_cast<float4*>(c + 4) = __ldg(c4 + 1);
    *reinterpret_cast<float4*>(s) = __ldg(s4); *reinterpret_cast<float4*>(s + 4) = __ldg(s4 + 1);
    // backward: the inverse rotation (sin_sign = -1) and, on the q columns, the d^-1/2 of q = (W_q x + b_q) * d^-1/2 (HF:341);
    // the forward passes (+1, 1): both multiplications are exact
    const float sc = which == 0 ? q_scale : 1.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] *= sc; s[i] *= sin_sign * sc; }
    __nv_bfloat162* l2 = reinterpret_cast<__nv_bfloat162*>(&lo);
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = __bfloat1622float2(l2[i]), b = __bfloat1622float2(h2[i]);
        // x*cos + rotate_half(x)*sin : first half gets -x2*sin, second half gets +x1*sin
        const float a0 = a.x * c[2 * i] - b.x * s[2 * i], a1 = a.y * c[2 * i + 1] - b.y * s[2 * i + 1];
        const float b0 = b.x * c[2 * i] + a.x * s[2 * i], b1 = b.y * c[2 * i + 1] + a.y * s[2 * i + 1];
        l2[i] = __floats2bfloat162_rn(a0, a1);
        h2[i] = __floats2bfloat162_rn(b0, b1);
    }
    *reinterpret_cast<uint4*>(base) = lo;
    *reinterpret_cast<uint4*>(base + half) = hi;
}

// ------------------------------------------------------------------------------------------------
// Pooled read-outs over the encoder output (embed_text.py:112-129 masked mean; baselines/model.py:104-120 CLS)
// grid = (n_seq, ceil(h / 256)); thread = one channel; coalesced across channels.
// ------------------------------------------------------------------------------------------------
__global__ void pool_kernel(const __nv_bfloat16* __restrict__ enc, const int64_t* __restrict__ ids, int k_tokens, int h,
                            int mode, float* __restrict__ out) {
    const int n = blockIdx.x, c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= h) return;
    const __nv_bfloat16* base = enc + static_cast<size_t>(n) * k_tokens * h + c;
    if (mode == 1) { out[static_cast<size_t>(n) * h + c] = __bfloat162float(base[0]); return; }
    float acc = 0.f;
    int cnt = 0;
    for (int t = 0; t < k_tokens; ++t) {
        if (ids[static_cast<size_t>(n) * k_tokens + t] != 1) { acc += __bfloat162float(base[static_cast<size_t>(t) * h]); ++cnt; }
    }
    out[static_cast<size_t>(n) * h + c] = acc / fmaxf(static_cast<float>(cnt), 1e-9f);
}

}  // namespace

int embed_launch(const int64_t* ids, int n_seq, int k_tokens, const EmbedArgs& a, const void* word_emb,
                 const void* pos_emb, float* x, int32_t* kv_len, uint8_t* key_mask, int32_t* err_flag,
                 cudaStream_t stream) {
    MOLLY_CHECK(a.hidden % 8 == 0, MOLLY_ERR_UNSUPPORTED, "embed: hidden_size %% 8 != 0 (%d)", a.hidden);
    MOLLY_CHECK(a.position_type == 0 || pos_emb != nullptr, MOLLY_ERR_INVALID, "embed: absolute positions need pos_emb");
    int splits = (4 * device_sm_count() + n_seq - 1) / n_seq;      // ~4 CTAs per SM: a write-bound kernel needs the stores in flight
    splits = max(1, min(splits, (k_tokens + 63) / 64));
    const size_t smem = a.position_type == 1 ? sizeof(int32_t) * k_tokens : 0;
    MOLLY_CHECK(smem <= 48 * 1024, MOLLY_ERR_UNSUPPORTED, "embed: k_tokens %d too long for absolute positions", k_tokens);
    ProfScope prof(PF_EMBED, static_cast<double>(n_seq) * k_tokens * (a.hidden * 6.0 + 9.0), stream);
    embed_kernel<<<dim3(n_seq, splits), EMB_THREADS, smem, stream>>>(
        ids, k_tokens, a, static_cast<const __nv_bfloat16*>(word_emb), static_cast<const __nv_bfloat16*>(pos_emb), x,
        kv_len, key_mask, err_flag);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int mask_rows_launch(float* x, const uint8_t* key_mask, int rows, int h, cudaStream_t stream) {
    mask_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, key_mask, rows, h);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int layernorm_launch(const float* x, const float* w, const float* b, int rows, int h, float eps, void* out,
                     int out_dtype, cudaStream_t stream) {
    MOLLY_CHECK(h % 4 == 0 && h <= 20 * 128, MOLLY_ERR_UNSUPPORTED, "layernorm: h=%d unsupported (h %% 4, h <= 2560)", h);
    MOLLY_CHECK(rows > 0, MOLLY_ERR_INVALID, "layernorm: rows=%d", rows);
    ProfScope prof(PF_LAYERNORM, static_cast<double>(rows) * h * (out_dtype == DT_F32 ? 8.0 : 6.0), stream);
    if (out_dtype == DT_F32) launch_layernorm<float>(x, w, b, rows, h, eps, static_cast<float*>(out), stream);
    else launch_layernorm<__nv_bfloat16>(x, w, b, rows, h, eps, static_cast<__nv_bfloat16*>(out), stream);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int rotary_launch(void* qkv, int rows, int k_tokens, int h, int heads, const float* cos_t, const float* sin_t,
                  cudaStream_t stream, float sin_sign, float q_scale) {
    const int d = h / heads;
    MOLLY_CHECK(d % 16 == 0, MOLLY_ERR_UNSUPPORTED, "rotary: head_dim %d must be a multiple of 16", d);
    const long long items = static_cast<long long>(rows) * 2 * heads * (d / 16);
    ProfScope prof(PF_ROTARY, static_cast<double>(rows) * 2.0 * h * 4.0, stream);     // q,k read + write, bf16
    rotary_kernel<<<static_cast<unsigned>((items + 255) / 256), 256, 0, stream>>>(
        static_cast<__nv_bfloat16*>(qkv), items, k_tokens, h, heads, cos_t, sin_t, sin_sign, q_scale);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int pool_launch(const void* enc_out, const int64_t* ids, int n_seq, int k_tokens, int h, int mode, float* out,
                cudaStream_t stream) {
    pool_kernel<<<dim3(n_seq, (h + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(enc_out), ids,
                                                                   k_tokens, h, mode, out);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

}  // namespace molly
