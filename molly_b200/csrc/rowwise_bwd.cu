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
// forward kernel): one read of x and dy, one read-modify-write of d_x.
//   g = dy * gamma,  xhat = (x - mean) * rstd,  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
// d_x (fp32) receives dx (accumulate == 0) or has it added (accumulate != 0: the residual branch).  stats[row] = (mean, rstd).
template <int VPL>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const float* __restrict__ gamma, int rows,
              int h, float eps, float* __restrict__ d_x, int accumulate, float* __restrict__ stats) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, nvec = h >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * h);
    const uint2* dr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(row) * h);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    float4 xv[VPL], gv[VPL];                                   // x, then xhat;  g = dy * gamma
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        xv[i] = idx < nvec ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
    const float mean = wsum(s) / h;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
        if (lane + 32 * i < nvec) {
            const float a = xv[i].x - mean, b = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
    const float rstd = rsqrtf(wsum(ss) / h + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            const uint2 u = dr[idx];
            const float2 d0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 d1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            const float4 gm = __ldg(g4 + idx);
            gv[i] = make_float4(d0.x * gm.x, d0.y * gm.y, d1.x * gm.z, d1.y * gm.w);
            xv[i] = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd,
                                (xv[i].w - mean) * rstd);
            sg += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
            sgx += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
        } else {
            gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float mg = wsum(sg) / h, mgx = wsum(sgx) / h;
    float4* out = reinterpret_cast<float4*>(d_x + static_cast<size_t>(row) * h);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nvec) {
            float4 r = make_float4(rstd * (gv[i].x - mg - xv[i].x * mgx), rstd * (gv[i].y - mg - xv[i].y * mgx),
                                   rstd * (gv[i].z - mg - xv[i].z * mgx), rstd * (gv[i].w - mg - xv[i].w * mgx));
            if (accumulate) {
                const float4 o = out[idx];
                r = make_float4(r.x + o.x, r.y + o.y, r.z + o.z, r.w + o.w);
            }
            out[idx] = r;
        }
    }
    if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// d_gamma[c] += sum_rows dy * xhat, d_beta[c] += sum_rows dy: block = 32 columns x 32 row lanes over one chunk of rows,
// one atomicAdd per (chunk, column) into the pre-zeroed fp32 outputs.
__global__ void __launch_bounds__(1024)
ln_param_grad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const float* __restrict__ stats,
                     int rows, int h, int rows_per_chunk, float* __restrict__ d_gamma, float* __restrict__ d_beta) {
    __shared__ float sg[32][33], sb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    float ag = 0.f, ab = 0.f;
    if (c < h) {
        for (int r = r0 + ty; r < r1; r += 32) {
            const float d = __bfloat162float(dy[static_cast<size_t>(r) * h + c]);
            ag += d * (x[static_cast<size_t>(r) * h + c] - stats[2 * r]) * stats[2 * r + 1];
            ab += d;
        }
    }
    sg[ty][tx] = ag;
    sb[ty][tx] = ab;
    __syncthreads();
    if (ty == 0 && c < h) {
        float g = 0.f, b = 0.f;
        for (int i = 0; i < 32; ++i) { g += sg[i][tx]; b += sb[i][tx]; }
        atomicAdd(d_gamma + c, g);
        atomicAdd(d_beta + c, b);
    }
}

// act = gelu_erf(pre), d_pre = d_act * (Phi(pre) + pre * phi(pre))          (HF:57-61 exact-erf GELU); 8 elements per thread.
// erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7 on the exact-erf GELU, far below the bf16 rounding of the outputs); its
// exp(-x^2/2) factor is also the pdf of the derivative, so one ex2 serves both.
__global__ void __launch_bounds__(256)
gelu_fwd_bwd_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ d_act, long long n,
                    __nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ d_pre) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    const uint4 pv = *reinterpret_cast<const uint4*>(pre + i);
    uint4 dv = make_uint4(0, 0, 0, 0);
    if (d_act != nullptr) dv = *reinterpret_cast<const uint4*>(d_act + i);
    const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(&pv);
    const __nv_bfloat162* dp = reinterpret_cast<const __nv_bfloat162*>(&dv);
    uint4 av, gv;
    uint32_t* ap = reinterpret_cast<uint32_t*>(&av);
    uint32_t* gp = reinterpret_cast<uint32_t*>(&gv);
    auto f = [](float x, float dd, float& a, float& g) {
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
    };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 p = __bfloat1622float2(pp[k]), d = __bfloat1622float2(dp[k]);
        float a0, a1, g0, g1;
        f(p.x, d.x, a0, g0);
        f(p.y, d.y, a1, g1);
        ap[k] = pack_bf16x2(a0, a1);
        gp[k] = pack_bf16x2(g0, g1);
    }
    *reinterpret_cast<uint4*>(act + i) = av;
    if (d_act != nullptr) *reinterpret_cast<uint4*>(d_pre + i) = gv;
}

// u = (a_0, b_0, a_1, b_1, ...) interleaved [rows, 2F]; act = silu(a) * b [rows, F]; d_u = (d_a, d_b) interleaved
__global__ void __launch_bounds__(256)
glu_fwd_bwd_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ d_act, long long n_out,
                   __nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ d_u) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const float2 ab = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(u + 2 * i));
    const float d = d_act != nullptr ? __bfloat162float(d_act[i]) : 0.f;
    const float sig = 1.0f / (1.0f + __expf(-ab.x));
    const float sl = ab.x * sig;
    act[i] = __float2bfloat16(sl * ab.y);
    const float da = d * ab.y * sig * (1.0f + ab.x * (1.0f - sig));
    const float db = d * sl;
    if (d_act != nullptr) *reinterpret_cast<__nv_bfloat162*>(d_u + 2 * i) = __floats2bfloat162_rn(da, db);
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
                  int accumulate, float* stats, float* d_gamma, float* d_beta, cudaStream_t stream) {
    MOLLY_CHECK(rows > 0 && h > 0 && h % 4 == 0 && h <= 20 * 128, MOLLY_ERR_UNSUPPORTED,
                "ln_bwd: rows=%d h=%d (h %% 4 == 0, h <= 2560)", rows, h);
    const auto* dy = static_cast<const __nv_bfloat16*>(dy_bf16);
    ProfScope prof(PF_ROWWISE_BWD, static_cast<double>(rows) * h * (d_gamma ? 20.0 : 14.0), stream);
    {
        const int vpl = (h / 4 + 31) / 32;
#define MOLLY_LNB_CASE(V) \
        if (vpl <= V) ln_bwd_kernel<V><<<(rows + 7) / 8, 256, 0, stream>>>(x, dy, gamma, rows, h, eps, d_x, accumulate, stats); else
        MOLLY_LNB_CASE(1) MOLLY_LNB_CASE(2) MOLLY_LNB_CASE(4) MOLLY_LNB_CASE(8) MOLLY_LNB_CASE(10) MOLLY_LNB_CASE(20) {}
#undef MOLLY_LNB_CASE
    }
    count_launch();
    if (d_gamma != nullptr && d_beta != nullptr) {
        const int chunks = max(1, min(64, rows / 64));
        const int rpc = (rows + chunks - 1) / chunks;
        ln_param_grad_kernel<<<dim3((h + 31) / 32, chunks), 1024, 0, stream>>>(x, dy, stats, rows, h, rpc, d_gamma, d_beta);
        count_launch();
    }
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int act_fwd_bwd_launch(int glu, const void* pre, const void* d_act, long long rows, int f_out, void* act, void* d_pre,
                       cudaStream_t stream) {
    const long long n = rows * f_out;
    MOLLY_CHECK(n > 0 && f_out % 8 == 0, MOLLY_ERR_INVALID, "act_fwd_bwd: rows=%lld F=%d (F must be a multiple of 8)", rows,
                f_out);
    ProfScope prof(PF_ROWWISE_BWD, static_cast<double>(n) * (glu ? 12.0 : 8.0), stream);
    if (glu)
        glu_fwd_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(pre), static_cast<const __nv_bfloat16*>(d_act), n,
            static_cast<__nv_bfloat16*>(act), static_cast<__nv_bfloat16*>(d_pre));
    else
        gelu_fwd_bwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(pre), static_cast<const __nv_bfloat16*>(d_act), n,
            static_cast<__nv_bfloat16*>(act), static_cast<__nv_bfloat16*>(d_pre));
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
