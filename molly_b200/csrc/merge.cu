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

// =================================================================================================
// SURVEY.md 8f row N1 -- the input producer on the device: the placeholder runs found in input_ids are the index
// source (omic_info_list becomes a cross-check), and the LLM embedding lookup skips every row the path overwrites.
// =================================================================================================
namespace molly {
namespace {

// One CTA per sample.  For every position: j = index inside its run of *_pad tokens (or -1), and per run (in text
// order): first position, kind, length.  Two block-wide scans per 256-token chunk: a prefix MAX of "last non-pad
// position" (warp ballot + clz) and a prefix COUNT of run starts (ballot + popc), both with a running carry.
__global__ void __launch_bounds__(SCAN_THREADS)
placeholder_runs_kernel(const int64_t* __restrict__ input_ids, int T, int64_t pad0, int64_t pad1, int64_t pad2,
                        const int32_t* __restrict__ n_slots, int max_runs, int32_t* __restrict__ run_start, int32_t* __restrict__ run_kind,
                        int32_t* __restrict__ run_len, int32_t* __restrict__ n_runs, int32_t* __restrict__ pos_j) {
    __shared__ int s_last[SCAN_THREADS / 32];
    __shared__ int s_cnt[SCAN_THREADS / 32];
    __shared__ int s_carry_last, s_carry_cnt;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t* row = input_ids + static_cast<size_t>(b) * T;
    if (threadIdx.x == 0) { s_carry_last = -1; s_carry_cnt = 0; }
    for (int r = threadIdx.x; r < max_runs; r += SCAN_THREADS) run_len[static_cast<size_t>(b) * max_runs + r] = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += SCAN_THREADS) {
        const int t = t0 + threadIdx.x;
        int kind = -1;
        if (t < T) {
            const int64_t id = row[t];
            kind = (id == pad0) ? 0 : (id == pad1) ? 1 : (id == pad2) ? 2 : -1;
        }
        const bool pad = kind >= 0;
        // last non-pad position <= t inside the warp (positions beyond T count as non-pad: they end a run)
        const unsigned nonpad = __ballot_sync(0xffffffffu, !pad);
        const unsigned below = nonpad & (0xffffffffu >> (31 - lane));
        int last = below ? (t0 + warp * 32 + 31 - __clz(below)) : -2;        // -2: none in this warp so far
        if (lane == 31) s_last[warp] = nonpad ? (t0 + warp * 32 + 31 - __clz(nonpad)) : -2;
        __syncthreads();
        if (last == -2) {
            last = s_carry_last;
            for (int w = 0; w < warp; ++w) if (s_last[w] != -2) last = s_last[w];
        }
        const int j = pad ? t - last - 1 : -1;
        const bool is_start = pad && j == 0;
        const unsigned sb = __ballot_sync(0xffffffffu, is_start);
        const int rank = __popc(sb & ((1u << lane) - 1u));
        if (lane == 0) s_cnt[warp] = __popc(sb);
        __syncthreads();
        int before = s_carry_cnt;
        for (int w = 0; w < warp; ++w) before += s_cnt[w];
        // runs beyond the sample's omic_ids slots are never overwritten (the reference's zip stops): plain text rows
        const int my_run = before + __popc(sb & (0xffffffffu >> (31 - lane))) - 1;
        if (t < T) pos_j[static_cast<size_t>(b) * T + t] = (pad && n_slots && my_run >= n_slots[b]) ? -1 : j;
        if (is_start && before + rank < max_runs) {
            run_start[static_cast<size_t>(b) * max_runs + before + rank] = t;
            run_kind[static_cast<size_t>(b) * max_runs + before + rank] = kind;
        }
        // run length: the LAST pad of a run (next position is non-pad or T) knows it: j + 1
        if (pad) {
            const bool next_pad = (t + 1 < T) && [&] { const int64_t nid = row[t + 1]; return nid == pad0 || nid == pad1 || nid == pad2; }();
            if (!next_pad) {
                const int r = my_run;
                if (r >= 0 && r < max_runs) run_len[static_cast<size_t>(b) * max_runs + r] = j + 1;
            }
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) {
            int l = s_carry_last;
            for (int w = 0; w < SCAN_THREADS / 32; ++w) if (s_last[w] != -2) l = s_last[w];
            s_carry_last = l;
            s_carry_cnt = before + __popc(sb);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) n_runs[b] = s_carry_cnt;
}

// inputs_embeds[b,t,:] = table[input_ids[b,t],:] for every row the omics path will NOT overwrite (one warp per row).
template <int ELEM_BYTES>
__global__ void __launch_bounds__(256)
embed_tokens_skip_kernel(const int64_t* __restrict__ input_ids, const int32_t* __restrict__ pos_j, int64_t pad0,
                         int64_t pad1, int cap_nt, int cap_pr, const uint8_t* __restrict__ table, int vocab, int D,
                         uint8_t* __restrict__ out, long long rows, int32_t* err_flag) {
    const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    int64_t id = input_ids[row];
    const int j = pos_j[row];
    if (j >= 0 && j < ((id == pad0 || id == pad1) ? cap_nt : cap_pr)) return;       // the projector epilogue writes it
    if (id < 0 || id >= vocab) { if (err_flag && lane == 0) atomicOr(err_flag, 16); id = 0; }   // MOLLY_ERRBIT_TOKEN
    const size_t row_bytes = static_cast<size_t>(D) * ELEM_BYTES;
    const uint4* s = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * row_bytes);
    uint4* d = reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * row_bytes);
    for (int v = lane; v < static_cast<int>(row_bytes / 16); v += 32) d[v] = __ldg(s + v);
}

// Runs that build_seq_table will REJECT (shorter than the K cap of their kind, kind != the modality of the omic_ids slot they
// pair with, or no slot at all) are never overwritten by a projector: their positions go back to pos_j = -1, so that the
// skipping lookup embeds them like any other token instead of leaving rows of the torch.empty output uninitialised.
// One CTA per sample; the runs of a sample are few.
__global__ void __launch_bounds__(256)
placeholder_reject_kernel(int32_t* __restrict__ pos_j, const int32_t* __restrict__ run_start,
                          const int32_t* __restrict__ run_kind, const int32_t* __restrict__ run_len,
                          const int32_t* __restrict__ n_runs, const int32_t* __restrict__ slot_expect, int T, int max_runs,
                          int cap_nt, int cap_pr) {
    const int b = blockIdx.x;
    const int nr = n_runs[b] < max_runs ? n_runs[b] : max_runs;
    for (int r = 0; r < nr; ++r) {
        const size_t o = static_cast<size_t>(b) * max_runs + r;
        const int kind = run_kind[o], len = run_len[o], expect = slot_expect[o];      // expect: 0 dna/rna, 1 protein, -1 no slot
        const bool is_pr = kind == 2;
        const bool ok = expect >= 0 && (expect == 1) == is_pr && len >= (is_pr ? cap_pr : cap_nt);
        if (ok) continue;
        const int s0 = run_start[o];
        for (int t = s0 + threadIdx.x; t < s0 + len && t < T; t += blockDim.x) pos_j[static_cast<size_t>(b) * T + t] = -1;
    }
}

// seq_table[n] = (b, start) with start = (first pad position of run `slot`) - 1, i.e. the x_start token: exactly
// info["start"] of the reference (omics_dataset.py:277).  Runs pair with omic_ids slots BY INDEX (the reference's zip).
__global__ void build_seq_table_kernel(const int32_t* __restrict__ b_idx, const int32_t* __restrict__ slot_idx, int n,
                                       const int32_t* __restrict__ run_start, const int32_t* __restrict__ run_kind,
                                       const int32_t* __restrict__ run_len, const int32_t* __restrict__ n_runs,
                                       int max_runs, int expect_protein, int k_need, int32_t* __restrict__ seq_table,
                                       int32_t* err_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = b_idx[i], slot = slot_idx[i];
    int start = -1;
    if (slot < n_runs[b] && slot < max_runs) {
        const size_t o = static_cast<size_t>(b) * max_runs + slot;
        const bool is_pr = run_kind[o] == 2;
        if (is_pr == (expect_protein != 0) && run_len[o] >= k_need) start = run_start[o] - 1;
    }
    if (start < 0 && err_flag) atomicOr(err_flag, 8);           // layout does not match the ids: MOLLY_ERRBIT_LAYOUT
    seq_table[2 * i] = b;
    seq_table[2 * i + 1] = start;
}

}  // namespace

int placeholder_runs_launch(const int64_t* input_ids, int B, int T, int64_t pad0, int64_t pad1, int64_t pad2,
                            const int32_t* n_slots, int max_runs, int32_t* run_start, int32_t* run_kind, int32_t* run_len, int32_t* n_runs,
                            int32_t* pos_j, cudaStream_t stream) {
    MOLLY_CHECK(B > 0 && T > 0 && max_runs > 0, MOLLY_ERR_INVALID, "placeholder_runs: B=%d T=%d max_runs=%d", B, T, max_runs);
    placeholder_runs_kernel<<<B, SCAN_THREADS, 0, stream>>>(input_ids, T, pad0, pad1, pad2, n_slots, max_runs, run_start,
                                                           run_kind, run_len, n_runs, pos_j);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int placeholder_reject_launch(int32_t* pos_j, const int32_t* run_start, const int32_t* run_kind, const int32_t* run_len,
                              const int32_t* n_runs, const int32_t* slot_expect, int B, int T, int max_runs, int cap_nt,
                              int cap_pr, cudaStream_t stream) {
    MOLLY_CHECK(B > 0 && T > 0 && max_runs > 0, MOLLY_ERR_INVALID, "placeholder_reject: B=%d T=%d max_runs=%d", B, T, max_runs);
    placeholder_reject_kernel<<<B, 256, 0, stream>>>(pos_j, run_start, run_kind, run_len, n_runs, slot_expect, T, max_runs,
                                                     cap_nt, cap_pr);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int embed_tokens_skip_launch(const int64_t* input_ids, const int32_t* pos_j, int64_t pad0, int64_t pad1, int cap_nt,
                             int cap_pr, const void* table, int dtype, int vocab, int D, void* out, int B, int T,
                             int32_t* err_flag, cudaStream_t stream) {
    const int eb = dtype == DT_F32 ? 4 : 2;
    MOLLY_CHECK((static_cast<long long>(D) * eb) % 16 == 0, MOLLY_ERR_UNSUPPORTED, "embed_tokens: row bytes must be a 16-B multiple");
    const long long rows = static_cast<long long>(B) * T;
    const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
    ProfScope prof(PF_EMBED, static_cast<double>(rows) * D * eb * 2.0, stream);
    if (eb == 4)
        embed_tokens_skip_kernel<4><<<grid, 256, 0, stream>>>(input_ids, pos_j, pad0, pad1, cap_nt, cap_pr,
                                                              static_cast<const uint8_t*>(table), vocab, D,
                                                              static_cast<uint8_t*>(out), rows, err_flag);
    else
        embed_tokens_skip_kernel<2><<<grid, 256, 0, stream>>>(input_ids, pos_j, pad0, pad1, cap_nt, cap_pr,
                                                              static_cast<const uint8_t*>(table), vocab, D,
                                                              static_cast<uint8_t*>(out), rows, err_flag);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

int build_seq_table_launch(const int32_t* b_idx, const int32_t* slot_idx, int n, const int32_t* run_start,
                           const int32_t* run_kind, const int32_t* run_len, const int32_t* n_runs, int max_runs,
                           int expect_protein, int k_need, int32_t* seq_table, int32_t* err_flag, cudaStream_t stream) {
    MOLLY_CHECK(n > 0, MOLLY_ERR_INVALID, "build_seq_table: n=%d", n);
    build_seq_table_kernel<<<(n + 127) / 128, 128, 0, stream>>>(b_idx, slot_idx, n, run_start, run_kind, run_len, n_runs,
                                                               max_runs, expect_protein, k_need, seq_table, err_flag);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

}  // namespace molly
