// Microbenchmark: per-SM issue throughput of the instructions the attention softmax is made of (sm_100a).
// Measured on B200 (round 1): MUFU.EX2 16.0 /clk/SM; EX2 + F2FP interleaved 24.0 instr/clk/SM (the packs ride along for free);
// FFMA 126, FFMA2 63.9 instr/clk/SM (= 128 FMA lanes either way), FMNMX 128, integer pack (2 IADD + PRMT) 32.4 pairs/clk/SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int UNROLL = 16;

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* clk, float seed) {
    float r[UNROLL];
    uint32_t p[UNROLL / 2];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) r[i] = seed + threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < UNROLL / 2; ++i) p[i] = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0 || MODE == 2) {               // MUFU.EX2
#pragma unroll
            for (int i = 0; i < UNROLL; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
        }
        if (MODE == 1 || MODE == 2) {               // F2FP.BF16.F32.PACK_AB
#pragma unroll
            for (int i = 0; i < UNROLL / 2; ++i) {
                uint32_t q;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q) : "f"(r[2 * i]), "f"(r[2 * i + 1]));
                p[i] ^= q;
            }
        }
        if (MODE == 3) {                            // FFMA
#pragma unroll
            for (int i = 0; i < UNROLL; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(r[i]) : "f"(seed));
        }
        if (MODE == 4) {                            // FFMA2
#pragma unroll
            for (int i = 0; i < UNROLL / 2; ++i) {
                uint64_t a, b;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(r[2 * i]), "f"(r[2 * i + 1]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(seed));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a) : "l"(b));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(r[2 * i]), "=f"(r[2 * i + 1]) : "l"(a));
            }
        }
        if (MODE == 5) {                            // FMNMX (max)
#pragma unroll
            for (int i = 0; i < UNROLL; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(seed));
        }
        if (MODE == 6) {                            // integer pack: 2 IADD + PRMT per pair
#pragma unroll
            for (int i = 0; i < UNROLL / 2; ++i) {
                uint32_t a = __float_as_uint(r[2 * i]) + 0x8000u, b = __float_as_uint(r[2 * i + 1]) + 0x8000u, q;
                asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(q) : "r"(a), "r"(b));
                p[i] ^= q;
                r[2 * i] = __uint_as_float(q);
            }
        }
        if (MODE == 7) {                            // F2FP with a shared-memory store of the result (STS.128 every 4)
#pragma unroll
            for (int i = 0; i < UNROLL / 2; ++i) {
                uint32_t q;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q) : "f"(r[2 * i]), "f"(r[2 * i + 1]));
                p[i] = q;
            }
            __shared__ uint4 buf[512 * 2];
            buf[threadIdx.x * 2] = make_uint4(p[0], p[1], p[2], p[3]);
            buf[threadIdx.x * 2 + 1] = make_uint4(p[4], p[5], p[6], p[7]);
        }
    }
    const long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) acc += r[i];
#pragma unroll
    for (int i = 0; i < UNROLL / 2; ++i) acc += __uint_as_float(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread) {
    float* out;
    long long* clk;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&clk, 148 * 8);
    k<MODE><<<148, 512>>>(out, clk, 0.5f);
    k<MODE><<<148, 512>>>(out, clk, 0.5f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    printf("%-34s %8.2f thread-ops/clk/SM   (%.0f clk)\n", name, ops_per_iter_per_thread * ITERS * 512 / c, c);
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    run<0>("MUFU.EX2", UNROLL);
    run<2>("EX2 + F2FP interleaved (instrs)", UNROLL + UNROLL / 2);
    run<3>("FFMA", UNROLL);
    run<4>("FFMA2 (per instr)", UNROLL / 2);
    run<5>("FMNMX", UNROLL);
    run<6>("int pack 2xIADD+PRMT (per pair)", UNROLL / 2);
    return 0;
}
