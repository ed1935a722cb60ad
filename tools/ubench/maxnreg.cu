// Microbenchmark: does `setmaxnreg` change the issue rate of a MUFU.EX2 stream?  512 threads per CTA (128 registers at launch);
// warps 0-3 (one per scheduler) time 64 passes of 128 MUFU.EX2 (+ packs) per thread, warps 8-15 give registers away
// (setmaxnreg.dec 56) and exit, warps 4-7 idle.  MODE 0: no setmaxnreg, 1: warps 0-7 setmaxnreg.inc 200, 2: inc 200 AND the
// loop's operands live in registers R128+ (a 128-float array kept alive next to it).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o maxnreg maxnreg.cu && ./maxnreg
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk, const float* in) {
    const int warp = threadIdx.x >> 5;
    if (MODE != 0) {
        if (warp >= 8) { asm volatile("setmaxnreg.dec.sync.aligned.u32 56;"); }
        else { asm volatile("setmaxnreg.inc.sync.aligned.u32 200;"); }
    }
    float acc = 0.f;
    if (warp < 4) {
        float s[128], keep[64];
#pragma unroll
        for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x + i * 32) & 1023];
#pragma unroll
        for (int i = 0; i < 64; ++i) keep[i] = in[(threadIdx.x + i * 7) & 1023];
        const long long t0 = clock64();
        for (int it = 0; it < 64; ++it) {
            uint32_t pk = 0;
#pragma unroll
            for (int p = 0; p < 64; ++p) {
                s[2 * p] = ex2(s[2 * p]);
                s[2 * p + 1] = ex2(s[2 * p + 1]);
            }
#pragma unroll
            for (int p = 0; p < 64; ++p) {
                uint32_t q;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q) : "f"(s[2 * p + 1]), "f"(s[2 * p]));
                pk ^= q;
                s[2 * p] = s[2 * p] * 0.5f - 1.0f; s[2 * p + 1] = s[2 * p + 1] * 0.5f - 1.0f;
            }
            acc += __uint_as_float(pk);
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 64; ++i) keep[i] += acc;      // keeps 64 more registers live across the loop
            }
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
#pragma unroll
        for (int i = 0; i < 128; ++i) acc += s[i];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc += keep[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name) {
    float *out, *in;
    long long* clk;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&in, 2048 * 4);
    cudaMemset(in, 0, 2048 * 4);
    cudaMalloc(&clk, 148 * 8);
    k<MODE><<<148, 512>>>(out, clk, in);
    k<MODE><<<148, 512>>>(out, clk, in);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    printf("%-44s %7.1f clk per 128 MUFU.EX2 per warp (%.2f each)  [%s]\n", name, c / 64, c / 64 / 128, cudaGetErrorString(e));
    cudaFree(out); cudaFree(in); cudaFree(clk);
}

int main() {
    run<0>("no setmaxnreg (128 registers)");
    run<1>("setmaxnreg.inc 200");
    run<2>("setmaxnreg.inc 200, 64 more live registers");
    return 0;
}
