// Row-wise backward kernels of the encoder: LayerNorm, erf-GELU, gated SiLU, casts / scales and the
// embedding scatter-add.  All HBM-bound; plain coalesced CUDA, fp32 math.
#include <math_constants.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LayerNorm backward w.r.t. its input, one warp per row, the row held in registers (VPL float4 vectors per lane, like the
// forward kernel).  Every load of the row (x, dy and -- when accumulating -- the old d_x) is issued before the first
// reduction, so one row keeps ~25 B per column in flight per lane instead of three dependent memory phases.
//   g = dy * gamma,  xhat = (x - mean) * rstd,  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
// d_x (fp32) receives dx (accumulate == 0) or has it added (accumulate != 0: the residual branch); dy_next (optional) gets
// the bf16 copy of the NEW d_x: the output gradient of the linear layer below, which the dgrad / wgrad GEMMs read.
// stats[row] = (mean, rstd) for ln_colstats_kernel.
// Wide rows (VPL >= 10: ~150 registers) run 128-thread blocks so that three of them share an SM (12 warps instead of 8).
template <int VPL>
__global__ void __launch_bounds__(VPL >= 10 ? 128 : 256)
ln_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const float* __restrict__ gamma, int rows,
              int h, float eps, float* __restrict__ d_x, int accumulate, float* __restrict__ stats,
              __nv_bfloat16* __restrict__ dy_next) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, nvec = h >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * h);
    const uint2* dr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(row) * h);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    float4* out = reinterpret_cast<float4*>(d_x + static_cast<size_t>(row) * h);
    float4 xv[VPL], ov[VPL];                                   // x, then xhat;  old d_x
    uint2 dv[VPL];                                             // dy stays packed: g = dy * gamma is formed twice, not kept
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        const bool in = idx < nvec;
        xv[i] = in ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        dv[i] = in ? dr[idx] : make_uint2(0u, 0u);
        ov[i] = (in && accumulate) ? out[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    const float mean = wsum(s) / h;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
        if (lane + 32 * i < nvec) {
            const float a = xv[i].x - mean, b = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
    const float rstd = rsqrtf(wsum(ss) / h + eps);
    auto g_of = [&](int i, int idx) {
        const float2 d0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv[i].x));
        const float2 d1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv[i].y));
        const float4 gm = __ldg(g4 + idx);
        return make_float4(d0.x * gm.x, d0.y * gm.y, d1.x * gm.z, d1.y * gm.w);
    };
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            const float4 g = g_of(i, idx);
            xv[i] = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd,
                                (xv[i].w - mean) * rstd);
            sg += (g.x + g.y) + (g.z + g.w);
            sgx += (g.x * xv[i].x + g.y * xv[i].y) + (g.z * xv[i].z + g.w * xv[i].w);
        }
    }
    const float mg = wsum(sg) / h, mgx = wsum(sgx) / h;
    uint2* nxt = dy_next != nullptr ? reinterpret_cast<uint2*>(dy_next + static_cast<size_t>(row) * h) : nullptr;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            const float4 g = g_of(i, idx);
            const float4 r = make_float4(rstd * (g.x - mg - xv[i].x * mgx) + ov[i].x, rstd * (g.y - mg - xv[i].y * mgx) + ov[i].y,
                                         rstd * (g.z - mg - xv[i].z * mgx) + ov[i].z, rstd * (g.w - mg - xv[i].w * mgx) + ov[i].w);
            out[idx] = r;
            if (nxt != nullptr) nxt[idx] = make_uint2(pack_bf16x2(r.x, r.y), pack_bf16x2(r.z, r.w));
        }
    }
    if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// Column sums of one LayerNorm backward, one pass over x, dy and dy_next:
//   d_gamma[c] += sum_r dy * xhat,   d_beta[c] += sum_r dy,   d_bias[c] += sum_r dy_next  (the bias of the linear layer whose
//   output gradient dy_next is; optional)
// A thread owns 4 columns (one 16-B load of x, one 8-B load of dy / dy_next per row) and walks every 8th row of its chunk; the
// 8 row lanes of a block meet in shared memory, one atomicAdd per (chunk, column) into pre-zeroed fp32.
__global__ void __launch_bounds__(256)
ln_colstats_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const float* __restrict__ stats,
                   const __nv_bfloat16* __restrict__ dy_next, int rows, int h, int rows_per_chunk, float* __restrict__ d_gamma,
                   float* __restrict__ d_beta, float* __restrict__ d_bias) {
    __shared__ float sm[3][8][128 + 4];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 128 + tx * 4;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    const bool bias = d_bias != nullptr;
    float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f}, an[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < h) {
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += 8) {
            const size_t o = static_cast<size_t>(r) * h + c;
            const float4 xv = *reinterpret_cast<const float4*>(x + o);
            const uint2 du = *reinterpret_cast<const uint2*>(dy + o);
            const float2 st = *reinterpret_cast<const float2*>(stats + 2 * r);
            const float2 d0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&du.x));
            const float2 d1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&du.y));
            ag[0] += d0.x * (xv.x - st.x) * st.y;
            ag[1] += d0.y * (xv.y - st.x) * st.y;
            ag[2] += d1.x * (xv.z - st.x) * st.y;
            ag[3] += d1.y * (xv.w - st.x) * st.y;
            ab[0] += d0.x; ab[1] += d0.y; ab[2] += d1.x; ab[3] += d1.y;
            if (bias) {
                const uint2 nu = *reinterpret_cast<const uint2*>(dy_next + o);
                const float2 n0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&nu.x));
                const float2 n1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&nu.y));
                an[0] += n0.x; an[1] += n0.y; an[2] += n1.x; an[3] += n1.y;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sm[0][ty][tx * 4 + k] = ag[k];
        sm[1][ty][tx * 4 + k] = ab[k];
        sm[2][ty][tx * 4 + k] = an[k];
    }
    __syncthreads();
    // 256 threads: threads 0..127 finish d_gamma and d_bias of column blockIdx.x*128 + t, threads 128..255 d_beta
    const int t = threadIdx.x & 127, which = threadIdx.x >> 7;
    const int col = blockIdx.x * 128 + t;
    if (col < h) {
        if (which == 0) {
            float g = 0.f, n = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { g += sm[0][i][t]; n += sm[2][i][t]; }
            atomicAdd(d_gamma + col, g);
            if (bias) atomicAdd(d_bias + col, n);
        } else {
            float b = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) b += sm[1][i][t];
            atomicAdd(d_beta + col, b);
        }
    }
}

// act = gelu_erf(pre), d_pre = d_act * (Phi(pre) + pre * phi(pre))          (HF:57-61 exact-erf GELU); 8 elements per thread.
// erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7 on the exact-erf GELU, far below the bf16 rounding of the outputs); its
// exp(-x^2/2) factor is also the pdf of the derivative, so one ex2 serves both.
__device__ __forceinline__ void gelu_one(float x, float dd, float& a, float& g) {
    const float z = x * 0.70710678118654752f, az = fabsf(z);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * az * az));      // exp(-x^2 / 2)
    const float cdf = 0.5f * (1.0f + copysignf(fmaf(-p, e, 1.0f), z));
    a = x * cdf;
    g = dd * fmaf(x * 0.3989422804014327f, e, cdf);
}

// 8 consecutive elements; colacc (optional): += the 8 d_pre values AS ROUNDED to bf16 (the bias gradient sums what the wgrad reads)
__device__ __forceinline__ void gelu_eight(const uint4& pv, const uint4& dv, uint4& av, uint4& gv, float* colacc) {
    const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(&pv);
    const __nv_bfloat162* dp = reinterpret_cast<const __nv_bfloat162*>(&dv);
    uint32_t* ap = reinterpret_cast<uint32_t*>(&av);
    uint32_t* gp = reinterpret_cast<uint32_t*>(&gv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 p = __bfloat1622float2(pp[k]), d = __bfloat1622float2(dp[k]);
        float a0, a1, g0, g1;
        gelu_one(p.x, d.x, a0, g0);
        gelu_one(p.y, d.y, a1, g1);
        ap[k] = pack_bf16x2(a0, a1);
        gp[k] = pack_bf16x2(g0, g1);
        if (colacc != nullptr) {
            const float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gp[k]));
            colacc[2 * k] += r.x;
            colacc[2 * k + 1] += r.y;
        }
    }
}

__global__ void __launch_bounds__(256)
gelu_fwd_bwd_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ d_act, long long n,
                    __nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ d_pre) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    const uint4 pv = *reinterpret_cast<const uint4*>(pre + i);
    uint4 dv = make_uint4(0, 0, 0, 0);
    if (d_act != nullptr) dv = *reinterpret_cast<const uint4*>(d_act + i);
    uint4 av, gv;
    gelu_eight(pv, dv, av, gv, nullptr);
    if (act != nullptr) *reinterpret_cast<uint4*>(act + i) = av;          // (the training tape may already hold act)
    if (d_act != nullptr) *reinterpret_cast<uint4*>(d_pre + i) = gv;
}

// The same with the bias gradient d_bias[c] += sum_rows d_pre[r, c] fused: a thread owns 8 columns and walks the rows of its
// chunk (grid = column blocks of 1024 x row chunks), so the column sums stay in registers; one atomicAdd per (chunk, column).
__global__ void __launch_bounds__(128)
gelu_fwd_bwd_colsum_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ d_act, int rows, int f,
                           int rows_per_chunk, __nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ d_pre,
                           float* __restrict__ d_bias) {
    const int c = (blockIdx.x * 128 + threadIdx.x) * 8;
    if (c >= f) return;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
        const size_t o = static_cast<size_t>(r) * f + c;
        const uint4 pv = *reinterpret_cast<const uint4*>(pre + o);
        const uint4 dv = *reinterpret_cast<const uint4*>(d_act + o);
        uint4 av, gv;
        gelu_eight(pv, dv, av, gv, acc);
        *reinterpret_cast<uint4*>(act + o) = av;
        *reinterpret_cast<uint4*>(d_pre + o) = gv;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(d_bias + c + k, acc[k]);
}

// u = (a_0, b_0, a_1, b_1, ...) interleaved [rows, 2F]; act = silu(a) * b [rows, F]; d_u = (d_a, d_b) interleaved.
// 8 outputs per thread: two 16-B loads of u, one of d_act, one 16-B store of act, two of d_u.
__global__ void __launch_bounds__(256)
glu_fwd_bwd_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ d_act, long long n_out,
                   __nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ d_u) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    if (i >= n_out) return;
    uint4 uv[2];
    uv[0] = *reinterpret_cast<const uint4*>(u + 2 * i);
    uv[1] = *reinterpret_cast<const uint4*>(u + 2 * i + 8);
    uint4 dv = make_uint4(0, 0, 0, 0);
    if (d_act != nullptr) dv = *reinterpret_cast<const uint4*>(d_act + i);
    const __nv_bfloat162* up = reinterpret_cast<const __nv_bfloat162*>(uv);
    const __nv_bfloat16* dp = reinterpret_cast<const __nv_bfloat16*>(&dv);
    uint4 av, gv[2];
    __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(&av);
    __nv_bfloat162* gp = reinterpret_cast<__nv_bfloat162*>(gv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float2 ab = __bfloat1622float2(up[k]);
        const float d = __bfloat162float(dp[k]);
        const float sig = 1.0f / (1.0f + __expf(-ab.x));
        const float sl = ab.x * sig;
        ap[k] = __float2bfloat16(sl * ab.y);
        gp[k] = __floats2bfloat162_rn(d * ab.y * sig * (1.0f + ab.x * (1.0f - sig)), d * sl);
    }
    if (act != nullptr) *reinterpret_cast<uint4*>(act + i) = av;
    if (d_act != nullptr) {
        *reinterpret_cast<uint4*>(d_u + 2 * i) = gv[0];
        *reinterpret_cast<uint4*>(d_u + 2 * i + 8) = gv[1];
    }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ in, long long n, __nv_bfloat16* __restrict__ out) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i) = o;
}

// x[r, 0:cols] *= scale  (bf16 [rows, ld], in place)
__global__ void __launch_bounds__(256)
scale_cols_kernel(__nv_bfloat16* __restrict__ x, int rows, int ld, int cols, float scale) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int cp = cols >> 1;
    if (i >= static_cast<long long>(rows) * cp) return;
    const int r = static_cast<int>(i / cp), c = static_cast<int>(i % cp) * 2;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(x + static_cast<size_t>(r) * ld + c);
    const float2 v = __bfloat1622float2(*p);
    *p = __floats2bfloat162_rn(v.x * scale, v.y * scale);
}

// table[index[r], :] += scale[r] * src[r, :]   (fp32 atomics; rows with scale 0 or a negative index are skipped)
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ index, const float* __restrict__ scale,
                        int rows, int h, float* __restrict__ table) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float s = scale[row];
    const int idx = index[row];
    if (s == 0.f || idx < 0) return;
    const float* sr = src + static_cast<size_t>(row) * h;
    float* tr = table + static_cast<size_t>(idx) * h;
    for (int c = threadIdx.x & 31; c < h; c += 32) atomicAdd(tr + c, s * sr[c]);
}

}  // namespace

int ln_bwd_launch(const float* x, const void* dy_bf16, const float* gamma, int rows, int h, float eps, float* d_x,
                  int accumulate, float* stats, float* d_gamma, float* d_beta, cudaStream_t stream, void* dy_next_bf16,
                  float* d_bias_next) {
    MOLLY_CHECK(rows > 0 && h > 0 && h % 4 == 0 && h <= 20 * 128, MOLLY_ERR_UNSUPPORTED,
                "ln_bwd: rows=%d h=%d (h %% 4 == 0, h <= 2560)", rows, h);
    MOLLY_CHECK(d_bias_next == nullptr || (dy_next_bf16 != nullptr && d_gamma != nullptr && d_beta != nullptr),
                MOLLY_ERR_INVALID, "ln_bwd: the bias column sums need dy_next and the d_gamma / d_beta pass");
    const auto* dy = static_cast<const __nv_bfloat16*>(dy_bf16);
    auto* dy_next = static_cast<__nv_bfloat16*>(dy_next_bf16);
    {
        // x + dy read, d_x written (+ read when accumulating), dy_next written
        ProfScope prof(PF_ROWWISE_BWD, static_cast<double>(rows) * h * (10.0 + (accumulate ? 4.0 : 0.0) + (dy_next ? 2.0 : 0.0)),
                       stream);
        const int vpl = (h / 4 + 31) / 32;
#define MOLLY_LNB_CASE(V) \
        if (vpl <= V) ln_bwd_kernel<V><<<(rows + (V >= 10 ? 3 : 7)) / (V >= 10 ? 4 : 8), V >= 10 ? 128 : 256, 0, stream>>>( \
            x, dy, gamma, rows, h, eps, d_x, accumulate, stats, dy_next); else
        MOLLY_LNB_CASE(1) MOLLY_LNB_CASE(2) MOLLY_LNB_CASE(4) MOLLY_LNB_CASE(8) MOLLY_LNB_CASE(10) MOLLY_LNB_CASE(20) {}
#undef MOLLY_LNB_CASE
    }
    count_launch();
    if (d_gamma != nullptr && d_beta != nullptr) {
        const int chunks = max(1, min(128, rows / 64));
        const int rpc = (rows + chunks - 1) / chunks;
        ProfScope prof(PF_ROWWISE_BWD, static_cast<double>(rows) * h * (d_bias_next ? 8.0 : 6.0), stream);
        ln_colstats_kernel<<<dim3((h + 127) / 128, chunks), 256, 0, stream>>>(x, dy, stats, d_bias_next ? dy_next : nullptr, rows,
                                                                             h, rpc, d_gamma, d_beta, d_bias_next);
        count_launch();
    }
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int act_fwd_bwd_launch(int glu, const void* pre, const void* d_act, long long rows, int f_out, void* act, void* d_pre,
                       cudaStream_t stream, float* d_bias) {
    const long long n = rows * f_out;
    MOLLY_CHECK(n > 0 && f_out % 8 == 0, MOLLY_ERR_INVALID, "act_fwd_bwd: rows=%lld F=%d (F must be a multiple of 8)", rows,
                f_out);
    MOLLY_CHECK(d_bias == nullptr || (!glu && d_act != nullptr && rows < (1LL << 31)), MOLLY_ERR_UNSUPPORTED,
                "act_fwd_bwd: the fused bias gradient covers the GELU backward only");
    ProfScope prof(PF_ROWWISE_BWD, static_cast<double>(n) * (glu ? 12.0 : 8.0), stream);
    if (glu) {
        glu_fwd_bwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(pre), static_cast<const __nv_bfloat16*>(d_act), n,
            static_cast<__nv_bfloat16*>(act), static_cast<__nv_bfloat16*>(d_pre));
    } else if (d_bias != nullptr) {           // pre-zeroed d_bias; ~4 resident blocks of 128 threads per SM, one wave
        const int col_blocks = (f_out + 1023) / 1024;
        int chunks = 4 * device_sm_count() / col_blocks;
        chunks = max(1, min(chunks, static_cast<int>(rows)));
        const int rpc = static_cast<int>((rows + chunks - 1) / chunks);
        gelu_fwd_bwd_colsum_kernel<<<dim3(col_blocks, (static_cast<int>(rows) + rpc - 1) / rpc), 128, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(pre), static_cast<const __nv_bfloat16*>(d_act), static_cast<int>(rows), f_out, rpc,
            static_cast<__nv_bfloat16*>(act), static_cast<__nv_bfloat16*>(d_pre), d_bias);
    } else {
        gelu_fwd_bwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(pre), static_cast<const __nv_bfloat16*>(d_act), n,
            static_cast<__nv_bfloat16*>(act), static_cast<__nv_bfloat16*>(d_pre));
    }
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int cast_f32_bf16_launch(const float* in, long long n, void* out, cudaStream_t stream) {
    MOLLY_CHECK(n > 0 && n % 4 == 0, MOLLY_ERR_INVALID, "cast: n=%lld must be a positive multiple of 4", n);
    ProfScope prof(PF_ROWWISE_BWD, static_cast<double>(n) * 6.0, stream);
    cast_f32_bf16_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, stream>>>(in, n,
                                                                                      static_cast<__nv_bfloat16*>(out));
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int scale_cols_launch(void* x_bf16, int rows, int ld, int cols, float scale, cudaStream_t stream) {
    MOLLY_CHECK(rows > 0 && cols > 0 && cols % 2 == 0 && cols <= ld, MOLLY_ERR_INVALID, "scale_cols: rows=%d cols=%d ld=%d",
                rows, cols, ld);
    const long long n = static_cast<long long>(rows) * (cols / 2);
    scale_cols_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(x_bf16), rows, ld,
                                                                                cols, scale);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int scatter_add_rows_launch(const float* src, const int32_t* index, const float* scale, int rows, int h, float* table,
                            cudaStream_t stream) {
    MOLLY_CHECK(rows > 0 && h > 0, MOLLY_ERR_INVALID, "scatter_add_rows: rows=%d h=%d", rows, h);
    scatter_add_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(src, index, scale, rows, h, table);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

}  // namespace molly
