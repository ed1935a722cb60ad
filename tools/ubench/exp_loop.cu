// Microbenchmark: cycles per element of the attention exp2 loop (FFMA2 -> 2x MUFU.EX2 -> FADD2, 128 elements in
// registers, as in softmax_block) as a function of the warps resident per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_loop exp_loop.cu
// Measured on B200 (round 1), cycles per 128-element pass per warp: 1 warp/scheduler 1 145 (MUFU 89 % busy), 2 warps 2 111, 3 warps 3 127.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>       // 0: full loop, 1: MUFU only, 2: loop + bf16 pack, 3: the kernel's chunked loop (FFMA2, 2 MUFU, FADD2, F2FP interleaved), 4: same with the integer-pipe pack
__global__ void k(float* out, long long* clk, const float* in, int iters) {
    float s[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x + i * 32) & 1023];
    float acc = 0.f;
    uint32_t pk = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint64_t sc2 = pack2(1.44f, 1.44f), nm2 = pack2(-acc * 1e-30f - 3.f, -acc * 1e-30f - 3.f);
        uint64_t sum2[2] = {pack2(0.f, 0.f), pack2(0.f, 0.f)};
        if (MODE == 3 || MODE == 4) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pkk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x0, x1;
                    unpack2(fma2(pack2(s[c * 32 + 2 * i], s[c * 32 + 2 * i + 1]), sc2, nm2), x0, x1);
                    x0 = ex2(x0); x1 = ex2(x1);
                    sum2[i & 1] = add2(sum2[i & 1], pack2(x0, x1));
                    if (MODE == 3) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pkk[i]) : "f"(x1), "f"(x0));
                    else pkk[i] = __byte_perm(__float_as_uint(x0) + 0x8000u, __float_as_uint(x1) + 0x8000u, 0x7632);
                    s[c * 32 + 2 * i] = x0 * 0.5f - 1.0f; s[c * 32 + 2 * i + 1] = x1 * 0.5f - 1.0f;
                }
                // stand-in for the tcgen05.st of the chunk: the 16 packed registers are consumed together
                asm volatile("" :: "r"(pkk[0]), "r"(pkk[1]), "r"(pkk[2]), "r"(pkk[3]), "r"(pkk[4]), "r"(pkk[5]), "r"(pkk[6]), "r"(pkk[7]),
                             "r"(pkk[8]), "r"(pkk[9]), "r"(pkk[10]), "r"(pkk[11]), "r"(pkk[12]), "r"(pkk[13]), "r"(pkk[14]), "r"(pkk[15]));
            }
        } else
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float x0, x1;
                if (MODE == 1) { x0 = s[i + 2 * u]; x1 = s[i + 2 * u + 1]; }
                else unpack2(fma2(pack2(s[i + 2 * u], s[i + 2 * u + 1]), sc2, nm2), x0, x1);
                s[i + 2 * u] = ex2(x0);
                s[i + 2 * u + 1] = ex2(x1);
                if (MODE != 1) sum2[u] = add2(sum2[u], pack2(s[i + 2 * u], s[i + 2 * u + 1]));
            }
        }
        if (MODE == 3 || MODE == 4) {
            // nothing here: MODE 3 / 4 replace the loop above (see below)
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
                uint32_t q;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q) : "f"(s[i]), "f"(s[i + 1]));
                pk ^= q;
            }
        }
        float a, b, c, d;
        unpack2(sum2[0], a, b);
        unpack2(sum2[1], c, d);
        acc += (a + b) + (c + d);
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 128; ++i) acc += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(pk);
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
    float *out, *in;
    long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&in, 1024 * 4);
    cudaMemset(in, 0, 1024 * 4);
    cudaMalloc(&clk, 148 * 8);
    const int iters = 64;
    k<MODE><<<148, warps_per_sm * 32>>>(out, clk, in, iters);
    k<MODE><<<148, warps_per_sm * 32>>>(out, clk, in, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    printf("%-22s warps/SMSP %d: %7.1f clk per 128-element pass per warp  (%.2f clk/element; MUFU lanes busy %.0f %%)\n", name,
           warps_per_sm / 4, c / iters, c / iters / 128, 100.0 * warps_per_sm * 32 * 128 * iters / (16.0 * c));
    cudaFree(out); cudaFree(in); cudaFree(clk);
}

int main() {
    for (int w : {4, 8, 12, 16}) run<1>("MUFU only", w);
    for (int w : {4, 8, 12, 16}) run<0>("FFMA2+MUFU+FADD2", w);
    for (int w : {4, 8, 12, 16}) run<2>("  ... + F2FP pack", w);
    for (int w : {4, 8, 12, 16}) run<3>("kernel loop (F2FP)", w);
    for (int w : {4, 8, 12, 16}) run<4>("kernel loop (int pack)", w);
    return 0;
}
