// Projector backward (training config, --train-mlp): dW[D,h] = dY^T X, db[D] = sum_rows dY, where dY are the rows of
// d(hidden_states) that the forward's slice-assign wrote (omics_one.py:91-97) and X is the saved encoder output.
// Both products contract over the token dimension, so dY and X are first transposed (smem-tiled, coalesced both ways)
// into K-major operands and the tcgen05 GEMM of gemm.cu does the contraction; db is a warp-per-row reduction of dY^T.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

namespace {

// out[c, r] = in[r, c]; in: [rows, cols] bf16, out: [cols, ld_out] bf16 (ld_out >= rows, padded columns zeroed)
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int rows, int cols, __nv_bfloat16* __restrict__ out,
                      int ld_out) {
    __shared__ __nv_bfloat16 tile[64][66];
    const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
    for (int i = ty; i < 64; i += 4) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? in[static_cast<size_t>(r) * cols + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 4) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols && r < ld_out) out[static_cast<size_t>(c) * ld_out + r] = tile[tx][i];
    }
}

__global__ void __launch_bounds__(256)
row_sum_kernel(const __nv_bfloat16* __restrict__ in, int rows, int cols, int ld, float* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int c = lane; c < cols; c += 32) acc += __bfloat162float(in[static_cast<size_t>(row) * ld + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = acc;
}

}  // namespace

// workspace: dyT [D, Mp] + xT [h, Mp] bf16, Mp = round_up(M, 8)
int project_bwd_launch(const void* dy_bf16, const void* x_bf16, int M, int D, int h, float* d_weight, float* d_bias,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int Mp = (M + 7) / 8 * 8;
    const size_t need = (static_cast<size_t>(D) + h) * Mp * 2;
    MOLLY_CHECK(workspace_bytes >= need, MOLLY_ERR_WORKSPACE, "project_bwd: workspace %zu < %zu", workspace_bytes, need);
    auto* dyT = static_cast<__nv_bfloat16*>(workspace);
    auto* xT = dyT + static_cast<size_t>(D) * Mp;
    transpose_bf16_kernel<<<dim3((Mp + 63) / 64, (D + 63) / 64), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(dy_bf16), M, D, dyT, Mp);
    count_launch();
    transpose_bf16_kernel<<<dim3((Mp + 63) / 64, (h + 63) / 64), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x_bf16), M, h, xT, Mp);
    count_launch();
    row_sum_kernel<<<(D + 7) / 8, 256, 0, stream>>>(dyT, D, M, Mp, d_bias);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    CUtensorMap ta, tb, tc;
    int rc = gemm_make_map_a(&ta, dyT, Mp, D, Mp);
    if (rc) return rc;
    if ((rc = gemm_make_map_b(&tb, xT, Mp, h, Mp, EPI_BIAS))) return rc;
    if ((rc = gemm_make_map_c(&tc, d_weight, DT_F32, h, D, h))) return rc;
    return gemm_launch(ta, tb, &tc, D, h, Mp, EPI_BIAS, nullptr, d_weight, DT_F32, h, nullptr, 0, 0, 0, 0, nullptr, stream);
}

// out[c, r] = in[r, c] (bf16): weight transposes for the dgrad GEMMs of the encoder backward
int transpose_bf16_launch(const void* in, int rows, int cols, void* out, cudaStream_t stream) {
    MOLLY_CHECK(rows > 0 && cols > 0, MOLLY_ERR_INVALID, "transpose: rows=%d cols=%d", rows, cols);
    ProfScope prof(PF_ROWWISE_BWD, static_cast<double>(rows) * cols * 4.0, stream);
    transpose_bf16_kernel<<<dim3((rows + 63) / 64, (cols + 63) / 64), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), rows, cols, static_cast<__nv_bfloat16*>(out), rows);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

}  // namespace molly
