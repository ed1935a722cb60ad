// Microbenchmark: how much does a second warp on the same scheduler slow the softmax exp2 loop (FFMA2, 2 x MUFU.EX2, FADD2,
// F2FP per key pair, 128 keys -- the loop of attention2.cu) when that second warp runs something ELSE?
// Warps 0-3 (one per scheduler) time 64 passes of the exp2 loop; warps 4-7 run a "noise" loop until the first group is done.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o interfere interfere.cu && ./interfere
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float max2(float a, float b) { float d; asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <int NOISE>      // 0 none (second group exits), 1 FMNMX3 stream, 2 FMNMX stream, 3 FFMA stream, 4 shared-memory loads, 5 mbarrier try_wait spin
__global__ void __launch_bounds__(256) k(float* out, long long* clk, const float* in) {
    __shared__ volatile int done;
    __shared__ float sm[256];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        done = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(1));
    }
    sm[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    float acc = 0.f;
    if (warp < 4) {
        float s[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x + i * 32) & 1023];
        const long long t0 = clock64();
        for (int it = 0; it < 64; ++it) {
            const uint64_t sc2 = pack2(1.44f, 1.44f), nm2 = pack2(-acc * 1e-30f - 3.f, -acc * 1e-30f - 3.f);
            uint64_t sum2[2] = {pack2(0.f, 0.f), pack2(0.f, 0.f)};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x0, x1;
                    unpack2(fma2(pack2(s[c * 32 + 2 * i], s[c * 32 + 2 * i + 1]), sc2, nm2), x0, x1);
                    x0 = ex2(x0); x1 = ex2(x1);
                    sum2[i & 1] = add2(sum2[i & 1], pack2(x0, x1));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(x1), "f"(x0));
                    s[c * 32 + 2 * i] = x0 * 0.5f - 1.0f; s[c * 32 + 2 * i + 1] = x1 * 0.5f - 1.0f;
                }
                asm volatile("" :: "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]),
                             "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15]));
            }
            float a, b, c, d;
            unpack2(sum2[0], a, b); unpack2(sum2[1], c, d);
            acc += (a + b) + (c + d);
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) { clk[blockIdx.x] = t1 - t0; }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) atomicAdd((int*)&done, 1);
#pragma unroll
        for (int i = 0; i < 128; ++i) acc += s[i];
    } else if (NOISE != 0) {
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = in[threadIdx.x + i];
        const uint32_t baddr = (uint32_t)__cvta_generic_to_shared(&bar);
        while (done < 4) {
#pragma unroll
            for (int rep = 0; rep < 16; ++rep) {
                if (NOISE == 1) { for (int i = 0; i < 8; ++i) m[i] = max3(m[i], m[(i + 1) & 7], m[(i + 3) & 7]); }
                if (NOISE == 2) { for (int i = 0; i < 8; ++i) m[i] = max2(m[i], m[(i + 1) & 7]); }
                if (NOISE == 3) { for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(m[i]) : "f"(m[(i + 1) & 7])); }
                if (NOISE == 4) { for (int i = 0; i < 8; ++i) m[i] += sm[(threadIdx.x + i * 32 + rep) & 255]; }
                if (NOISE == 5) {
                    uint32_t ok;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(baddr), "r"(0) : "memory");
                    m[0] += ok;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += m[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NOISE>
void run(const char* name) {
    float *out, *in;
    long long* clk;
    cudaMalloc(&out, 148 * 256 * 4);
    cudaMalloc(&in, 2048 * 4);
    cudaMemset(in, 0, 2048 * 4);
    cudaMalloc(&clk, 148 * 8);
    k<NOISE><<<148, 256>>>(out, clk, in);
    k<NOISE><<<148, 256>>>(out, clk, in);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    printf("second warp on the scheduler: %-28s exp2 loop %7.1f clk per 128 keys  (%.2f clk per MUFU.EX2)\n", name, c / 64, c / 64 / 128);
    cudaFree(out); cudaFree(in); cudaFree(clk);
}

int main() {
    run<0>("nothing");
    run<1>("FMNMX3 stream (3-input max)");
    run<2>("FMNMX stream (2-input max)");
    run<3>("FFMA stream");
    run<4>("LDS stream");
    run<5>("mbarrier.try_wait spin");
    return 0;
}
