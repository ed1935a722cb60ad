// Merge side of the path: placeholder scan over input_ids and the row scatter / gather between the projected omics
// rows and the LLM's [B, T, D] hidden_states (omics_one.py:93-97).  HBM-bound integer / copy kernels.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

namespace {

// ------------------------------------------------------------------------------------------------
// Placeholder scan: ordered compaction of the positions whose id is one of the three *_pad token ids.
// One CTA per sample; each 256-token chunk: warp ballot -> intra-warp rank by popc, warp totals -> smem prefix,
// running base carried across chunks.  Output order is ascending position (stable), exactly the order in which
// the dataset lays the runs out (omics_dataset.py:270-288).
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;

__global__ void __launch_bounds__(SCAN_THREADS)
placeholder_scan_kernel(const int64_t* __restrict__ input_ids, int T, int64_t pad0, int64_t pad1, int64_t pad2,
                        int32_t* __restrict__ out_pos, int32_t* __restrict__ out_kind, int32_t* __restrict__ out_counts) {
    __shared__ int s_warp[SCAN_THREADS / 32];
    __shared__ int s_base;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t* row = input_ids + static_cast<size_t>(b) * T;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += SCAN_THREADS) {
        const int t = t0 + threadIdx.x;
        int kind = -1;
        if (t < T) {
            const int64_t id = row[t];
            kind = (id == pad0) ? 0 : (id == pad1) ? 1 : (id == pad2) ? 2 : -1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, kind >= 0);
        const int rank = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (kind >= 0) {
            out_pos[static_cast<size_t>(b) * T + before + rank] = t;
            out_kind[static_cast<size_t>(b) * T + before + rank] = kind;
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) s_base = before + __popc(bal);
        __syncthreads();
    }
    if (threadIdx.x == 0) out_counts[b] = s_base;
}

// ------------------------------------------------------------------------------------------------
// Row scatter (un-fused form of the projector epilogue): hidden[b, start+1+j, :] = src[n*K + j, :]  for j < k_cap.
// One warp per source row, 16-B vectors.
// ------------------------------------------------------------------------------------------------
template <int ELEM_BYTES>
__global__ void __launch_bounds__(256)
merge_rows_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ seq_table, int rows, int k_tokens,
                  int k_cap, uint8_t* __restrict__ hidden, int B, int T, int D, int32_t* err_flag) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int n = row / k_tokens, j = row - n * k_tokens;
    const int b = __ldg(seq_table + 2 * n), start = __ldg(seq_table + 2 * n + 1);
    if (start < 0 || j >= k_cap) return;
    const int t = start + 1 + j;
    if (t >= T || b < 0 || b >= B) { if (err_flag && lane == 0) atomicOr(err_flag, 2); return; }
    const size_t row_bytes = static_cast<size_t>(D) * ELEM_BYTES;
    const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<size_t>(row) * row_bytes);
    uint4* d = reinterpret_cast<uint4*>(hidden + (static_cast<size_t>(b) * T + t) * row_bytes);
    const int nvec = static_cast<int>(row_bytes / 16);
    for (int v = lane; v < nvec; v += 32) d[v] = __ldg(s + v);
}

// ------------------------------------------------------------------------------------------------
// Backward of the slice-assign: gather dY rows (-> bf16 [rows, D]; skipped rows are zero) and zero them in d_hidden.
// ------------------------------------------------------------------------------------------------
template <typename T_>
__global__ void __launch_bounds__(256)
gather_grad_rows_kernel(T_* __restrict__ d_hidden, const int32_t* __restrict__ seq_table, int rows, int k_tokens,
                        int k_cap, int B, int T, int D, __nv_bfloat16* __restrict__ dy, int zero_rows) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int n = row / k_tokens, j = row - n * k_tokens;
    const int b = __ldg(seq_table + 2 * n), start = __ldg(seq_table + 2 * n + 1);
    __nv_bfloat16* o = dy + static_cast<size_t>(row) * D;
    const int t = start + 1 + j;
    const bool live = start >= 0 && j < k_cap && t < T && b >= 0 && b < B;
    T_* g = live ? d_hidden + (static_cast<size_t>(b) * T + t) * D : nullptr;
    for (int c = lane * 2; c < D; c += 64) {
        float v0 = 0.f, v1 = 0.f;
        if (live) {
            v0 = static_cast<float>(g[c]); v1 = static_cast<float>(g[c + 1]);
            if (zero_rows) { g[c] = static_cast<T_>(0.f); g[c + 1] = static_cast<T_>(0.f); }
        }
        *reinterpret_cast<__nv_bfloat162*>(o + c) = __floats2bfloat162_rn(v0, v1);
    }
}

}  // namespace

int placeholder_scan_launch(const int64_t* input_ids, int B, int T, int64_t pad0, int64_t pad1, int64_t pad2,
                            int32_t* out_pos, int32_t* out_kind, int32_t* out_counts, cudaStream_t stream) {
    MOLLY_CHECK(B > 0 && T > 0, MOLLY_ERR_INVALID, "placeholder_scan: B=%d T=%d", B, T);
    placeholder_scan_kernel<<<B, SCAN_THREADS, 0, stream>>>(input_ids, T, pad0, pad1, pad2, out_pos, out_kind,
                                                           out_counts);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int merge_rows_launch(const void* src, const int32_t* seq_table, int n_seq, int k_tokens, int k_cap, void* hidden,
                      int dtype, int B, int T, int D, int32_t* err_flag, cudaStream_t stream) {
    const int eb = dtype == DT_F32 ? 4 : 2;
    MOLLY_CHECK((static_cast<long long>(D) * eb) % 16 == 0, MOLLY_ERR_UNSUPPORTED, "merge: row bytes must be 16-B multiple");
    const int rows = n_seq * k_tokens;
    const int grid = (rows + 7) / 8;
    ProfScope prof(PF_MERGE, static_cast<double>(rows) * D * eb * 2.0, stream);
    if (eb == 4)
        merge_rows_kernel<4><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(src), seq_table, rows, k_tokens,
                                                       k_cap, static_cast<uint8_t*>(hidden), B, T, D, err_flag);
    else
        merge_rows_kernel<2><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(src), seq_table, rows, k_tokens,
                                                       k_cap, static_cast<uint8_t*>(hidden), B, T, D, err_flag);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int gather_grad_rows_launch(void* d_hidden, int dtype, const int32_t* seq_table, int n_seq, int k_tokens, int k_cap,
                            int B, int T, int D, void* dy_bf16, int zero_rows, cudaStream_t stream) {
    MOLLY_CHECK(D % 2 == 0, MOLLY_ERR_UNSUPPORTED, "gather_grad_rows: D must be even");
    const int rows = n_seq * k_tokens;
    const int grid = (rows + 7) / 8;
    if (dtype == DT_F32)
        gather_grad_rows_kernel<float><<<grid, 256, 0, stream>>>(static_cast<float*>(d_hidden), seq_table, rows,
                                                                 k_tokens, k_cap, B, T, D,
                                                                 static_cast<__nv_bfloat16*>(dy_bf16), zero_rows);
    else
        gather_grad_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
            static_cast<__nv_bfloat16*>(d_hidden), seq_table, rows, k_tokens, k_cap, B, T, D,
            static_cast<__nv_bfloat16*>(dy_bf16), zero_rows);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

}  // namespace molly
