// Projector backward (training config, --train-mlp): dW[D,h] = dY^T X, db[D] = sum_rows dY, where dY are the rows of
// d(hidden_states) that the forward's slice-assign wrote (omics_one.py:91-97) and X is the saved encoder output.
// Both products contract over the token dimension, so dY and X are first transposed (smem-tiled, coalesced both ways)
// into K-major operands and the tcgen05 GEMM of gemm.cu does the contraction; db is a warp-per-row reduction of dY^T.
#include <stdlib.h>

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


// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient without transposes: dW[N, K] = dY[M, N]^T X[M, K] contracts over the ROW dimension of both row-major
// operands, i.e. both are MN-major for the MMA (like V in the attention kernels).  TMA brings 64-row x 64-column boxes of
// dY and X (128-B swizzle); tcgen05.mma with a_major = b_major = MN; 128 x 128 fp32 accumulators (double-buffered in TMEM)
// are stored straight to dW.  d_bias = column sums of dY (colsum_bf16_kernel).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WG_TILE = 128;                   // output tile: 128 rows of dW (columns of dY) x 128 columns (columns of X)
constexpr int WG_BK = 64;                      // token rows per pipeline stage
constexpr int WG_STAGES = 6;
constexpr int WG_BOX_BYTES = WG_BK * 128;      // one 64 x 64 bf16 box
constexpr int WG_STAGE_BYTES = 4 * WG_BOX_BYTES;   // two boxes of dY + two boxes of X
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 256;
constexpr int WG_THREADS = 256;

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_mn_kernel(const __grid_constant__ CUtensorMap tma_dy, const __grid_constant__ CUtensorMap tma_x, int M, int N, int K,
                float* __restrict__ d_weight) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + WG_STAGES;
    uint64_t* tmem_full = empty_bar + WG_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_dy);
        tma_prefetch_desc(&tma_x);
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_k = (K + WG_TILE - 1) / WG_TILE, tiles_n = (N + WG_TILE - 1) / WG_TILE;
    const int num_tiles = tiles_n * tiles_k, num_mb = (M + WG_BK - 1) / WG_BK;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n0 = (tile / tiles_k) * WG_TILE, k0 = (tile % tiles_k) * WG_TILE;
                for (int mb = 0; mb < num_mb; ++mb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * WG_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
                    tma_load_2d(st, &tma_dy, &full_bar[stage], n0, mb * WG_BK);
                    tma_load_2d(st + WG_BOX_BYTES, &tma_dy, &full_bar[stage], n0 + 64, mb * WG_BK);
                    tma_load_2d(st + 2 * WG_BOX_BYTES, &tma_x, &full_bar[stage], k0, mb * WG_BK);
                    tma_load_2d(st + 3 * WG_BOX_BYTES, &tma_x, &full_bar[stage], k0 + 64, mb * WG_BK);
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(WG_TILE, WG_TILE, true, true);       // both operands MN-major
            int stage = 0, local = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
                const int acc = local & 1;
                mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * WG_TILE;
                for (int mb = 0; mb < num_mb; ++mb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES), sb = sa + 2 * WG_BOX_BYTES;
#pragma unroll
                    for (int s = 0; s < WG_BK / 16; ++s) {   // 16 token rows per MMA; LBO = next 64-wide atom, SBO = 8 rows
                        const uint64_t ad = make_smem_desc(sa + s * 16 * 128, WG_BOX_BYTES, 1024, kLayoutSW128);
                        const uint64_t bd = make_smem_desc(sb + s * 16 * 128, WG_BOX_BYTES, 1024, kLayoutSW128);
                        umma_bf16_ss(d_tmem, ad, bd, idesc, (mb | s) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;                             // TMEM lanes 32*ew .. +31 = rows of the dW tile
        const uint32_t lane_addr = static_cast<uint32_t>(ew * 32) << 16;
        int local = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            const int n0 = (tile / tiles_k) * WG_TILE, k0 = (tile % tiles_k) * WG_TILE;
            mbar_wait(&tmem_full[acc], (local >> 1) & 1);
            tc_fence_after();
            const int n = n0 + ew * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < WG_TILE; c += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + acc * WG_TILE + lane_addr + c, v);
                tmem_ld_wait();
                if (n < N) {
                    float* dst = d_weight + static_cast<size_t>(n) * K + k0 + c;
                    if (k0 + c + 32 <= K) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            reinterpret_cast<float4*>(dst)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                            __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                    } else {
                        for (int i = 0; i < 32; ++i)
                            if (k0 + c + i < K) dst[i] = __uint_as_float(v[i]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// out[c] += sum_r in[r, c]  (bf16 [rows, cols], cols % 8 == 0 -> fp32): a thread sums 8 columns (one 16-B load per row)
// over every 8th row of its chunk, the 8 row lanes of a block meet in shared memory, one atomicAdd per (chunk, column)
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ in, int rows, int cols, int rows_per_chunk, float* __restrict__ out) {
    __shared__ float sm[8][256 + 8];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 256 + tx * 8;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < cols) {
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(in + static_cast<size_t>(r) * cols + c);
            const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(p[k]);
                a[2 * k] += f.x;
                a[2 * k + 1] += f.y;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[ty][tx * 8 + k] = a[k];
    __syncthreads();
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
        atomicAdd(out + col, t);
    }
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

// dW[N, K] = dy[M, N]^T x[M, K], db[N] = colsum(dy), no transposes (wgrad_mn_kernel).  N and K must be multiples of 8.
int linear_wgrad_launch(const void* dy_bf16, const void* x_bf16, int M, int N, int K, float* d_weight, float* d_bias,
                        cudaStream_t stream, bool zeroed) {
    MOLLY_CHECK(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, MOLLY_ERR_UNSUPPORTED,
                "linear_wgrad: M=%d N=%d K=%d (N, K must be multiples of 8)", M, N, K);
    static const bool legacy = [] { const char* e = getenv("MOLLY_WGRAD_LEGACY"); return e != nullptr && e[0] == '1'; }();
    if (!legacy) {
        // the main tcgen05 GEMM with both operands MN-major (CTA pairs on 256 x 256 tiles, K split across the SMs)
        set_gemm_family(PF_GEMM_OTHER);
        int rc = gemm_launch_mn(GEMM_OPND_MN_MN, dy_bf16, N, x_bf16, K, N, K, M, d_weight, DT_F32, K, stream, zeroed);
        if (rc) return rc;
    } else {                                   // MOLLY_WGRAD_LEGACY=1: the first wgrad kernel (one CTA per 128 x 128 tile)
        CUtensorMap tdy, tx;
        int rc = make_tma_2d(&tdy, dy_bf16, M, N, N, WG_BK, 64, 2);
        if (rc) return rc;
        if ((rc = make_tma_2d(&tx, x_bf16, M, K, K, WG_BK, 64, 2))) return rc;
        static bool configured = false;
        if (!configured) {
            MOLLY_CUDA(cudaFuncSetAttribute(wgrad_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
            configured = true;
        }
        const int tiles = ((N + WG_TILE - 1) / WG_TILE) * ((K + WG_TILE - 1) / WG_TILE);
        const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
        {
            ProfScope prof(PF_GEMM_OTHER, 2.0 * M * N * static_cast<double>(K), stream);
            wgrad_mn_kernel<<<grid, WG_THREADS, WG_SMEM, stream>>>(tdy, tx, M, N, K, d_weight);
        }
        count_launch();
    }
    if (d_bias != nullptr) {
        if (!zeroed) MOLLY_CUDA(cudaMemsetAsync(d_bias, 0, sizeof(float) * N, stream));
        const int chunks = max(1, min(128, M / 64));
        const int rpc = (M + chunks - 1) / chunks;
        ProfScope prof(PF_ROWWISE_BWD, static_cast<double>(M) * N * 2.0, stream);
        colsum_bf16_kernel<<<dim3((N + 255) / 256, chunks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dy_bf16), M, N,
                                                                            rpc, d_bias);
        count_launch();
    }
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
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
