// bf16 x bf16 -> fp32 GEMM for sm_100a:  out[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
//   * operands move HBM -> shared memory with TMA (128-byte swizzle, 64-element K slabs), 3-6 stage mbarrier ring
//   * one elected thread issues tcgen05.mma (cta_group::1, M=128, N=BLOCK_N, K=16); accumulators live in TMEM,
//     double-buffered (2 x BLOCK_N fp32 columns) so the epilogue of tile i overlaps the main loop of tile i+1
//   * persistent: grid = #SMs, static round-robin over (m,n) tiles, n fastest so the CTAs that are co-resident
//     share A row-blocks through L2 while the (small) weight matrix stays L2 resident
//   * warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM alloc/dealloc, 4..11 = epilogue
//     (TMEM lane quarter = warp % 4, column half = (warp - 4) / 4)
//   * epilogue: tcgen05.ld (thread == accumulator row) -> fused math -> 128-B-swizzled smem staging (conflict free)
//     -> TMA store, so HBM sees full coalesced rows; the fp32 residual is brought in by TMA loads that are
//     double-buffered against the math.  Fused epilogues replace the reference's separate elementwise kernels:
//       bias (+ q *= d^-1/2)      (HF:329-341 q/k/v Linear; omics_one.py:91 projector)
//       bias + exact-erf GELU     (HF:406-414, 57-61)
//       bias + residual, fp32     (HF:365-375, 417-427)
//       gated SiLU                (NT-v2 FFN, weight rows interleaved at pack time)
//       bias + row scatter        (omics_one.py:91-97: projector output written straight into hidden_states)
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {


namespace {

#ifndef GEMM_ELECT
#define GEMM_ELECT 1
#endif
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                      // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int B_BOX_ROWS = 128;                  // rows of one TMA box of the weight map: every tile shape is built from these
constexpr int NUM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int STG_BUF_BYTES = 32 * 128;          // one staging buffer: 32 rows x 128 B (SWIZZLE_128B box)

constexpr int imin(int a, int b) { return a < b ? a : b; }

// CTA2: the tile belongs to a CTA PAIR (cluster of 2): 256 x BLOCK_N, each CTA stages its 128 A rows and HALF of the
// B rows; one tcgen05.mma.cta_group::2 (M=256) consumes both CTAs' shared memory and fills both CTAs' TMEM.
template <int BLOCK_N, int STG_BUFS_, bool CTA2 = false>
struct GemmCfg {
    static constexpr int B_STAGE_BYTES = (CTA2 ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;
    static constexpr int STG_BUFS = STG_BUFS_;      // 2: residual chunks are prefetched (short-K GEMMs); 1: deeper ring
    static constexpr int STG_BYTES = NUM_EPI_WARPS * STG_BUFS * STG_BUF_BYTES;  // 32 KB or 64 KB
    static constexpr int STAGES = imin(8, (227 * 1024 - 1024 - STG_BYTES) / (A_STAGE_BYTES + B_STAGE_BYTES));
    static constexpr int TMEM_COLS = 2 * BLOCK_N;           // 512 or 256: powers of two
    static constexpr int STG_OFFSET = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
    static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFFSET + 512;
    static_assert(STAGES >= 3, "not enough shared memory for a 3-stage ring");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

struct GemmParams {
    int M, N, K;
    const float* bias;
    void* out;                  // EPI_SCATTER only (direct stores); every other mode stores through tma_c
    int ldo;
    const int32_t* seq_table;   // EPI_SCATTER: [n_seq][2] = (b, start)
    int seq_k, B, T, k_cap;
    int32_t* err_flag;
    int scale_cols;             // EPI_BIAS / EPI_BIAS_ROPE: columns [0, scale_cols) are multiplied by `scale` after the
    float scale;                //           bias (q = (x Wq^T + bq) * d^-1/2, HF:341); scale_cols % 32 == 0
    const float* rope_cos_t;    // EPI_BIAS_ROPE: fp32 [head_dim/2][rope_len] cos / sin of t * 10000^(-2i/d), FREQUENCY-major
    const float* rope_sin_t;    //   so the 32 lanes of a warp (32 consecutive positions t) read 128 contiguous bytes
    int rope_len;
    int rope_cols;              //   columns [0, rope_cols) (= q and k) are rotated; rope_cols % 64 == 0
    int rope_head_dim;          //   16 | 32 | 64 ; position = row % seq_k  (row index inside the padded sequence, HF:103)
    int splits;                 // split-K: the K range is cut into `splits` work items per output tile, each adding its partial
                                //   tile into the (pre-zeroed) fp32 output with a TMA reduce-add; 1 = plain stores
    int reduce_out;             // 1: fp32 output tiles are ADDED to what `out` holds (EPI_BIAS_ACCUM) even without a K split
};

// NeoX rotary on one 64-column chunk held by one thread (its row): x*cos + rotate_half(x)*sin per head (HF:45-54).
// cos/sin come from fp32 tables built on the host exactly like HF:81-115; the tables are frequency-major so that the
// warp's 32 positions make one coalesced 128-B load per frequency.
template <int D>
__device__ __forceinline__ void rope_chunk64(float* v, int t, const float* __restrict__ cos_t,
                                             const float* __restrict__ sin_t, int rope_len) {
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
        const float cs = __ldg(cos_t + static_cast<size_t>(i) * rope_len + t);
        const float sn = __ldg(sin_t + static_cast<size_t>(i) * rope_len + t);
#pragma unroll
        for (int hl = 0; hl < 64 / D; ++hl) {
            const float x1 = v[hl * D + i], x2 = v[hl * D + i + D / 2];
            v[hl * D + i] = x1 * cs - x2 * sn;
            v[hl * D + i + D / 2] = x2 * cs + x1 * sn;
        }
    }
}

// 16-byte chunk c of staging row `lane` under the 128-B swizzle (matches CU_TENSOR_MAP_SWIZZLE_128B)
__device__ __forceinline__ uint4* stg_chunk(uint8_t* buf, int lane, int c) {
    return reinterpret_cast<uint4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4));
}

template <typename OutT>
__device__ __forceinline__ void store_direct(OutT* dst, const float* v);     // 32 values
template <>
__device__ __forceinline__ void store_direct<__nv_bfloat16>(__nv_bfloat16* dst, const float* v) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        d4[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
}
template <>
__device__ __forceinline__ void store_direct<float>(float* dst, const float* v) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

__device__ __forceinline__ void add_bias32(float* v, const float* bias, int col0) {
    const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 bb = __ldg(b4 + i);
        v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
    }
}

// OPND: which way the operands lie in memory.  GEMM_OPND_KK: A [M, K] and W [N, K] row-major, both K-major for the MMA (every
// forward GEMM).  GEMM_OPND_K_MN: the second operand is given as [K, N] row-major = MN-major (dgrad: d_in = d_out W with W as
// stored, no transpose).  GEMM_OPND_MN_MN: both operands are [K, M] / [K, N] row-major (wgrad: dW = dY^T X contracts over the
// rows of both).  An MN-major operand arrives as 64 x 64 boxes (64 K-rows x 128 B, 128-B swizzle), one box per 64 M/N values;
// the stage holds the same bytes, only the descriptors change (LBO = box pitch, SBO = 8 K-rows, 16 K-rows per MMA).
template <int BLOCK_N, int EPI, typename OutT, int STG_BUFS, bool CTA2, int OPND = GEMM_OPND_KK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r,
                    const GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, STG_BUFS, CTA2>;
    constexpr bool A_MN = OPND == GEMM_OPND_MN_MN, B_MN = OPND != GEMM_OPND_KK;
    constexpr int MN_BOX_BYTES = 64 * 128;
    constexpr int TILE_M = CTA2 ? 2 * BLOCK_M : BLOCK_M;
    const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs), 1 = peer
    const int tile_first = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tile_step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool kOutF32 = sizeof(OutT) == 4;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* res_bar = tmem_empty + 2;                      // [NUM_EPI_WARPS][2] residual-load barriers
    // the TMEM base lives in its own 16-B granule at the end of the barrier area: tcgen05.alloc's write is tracked by
    // compute-sanitizer racecheck as a 16-B access, which overlapped the neighbouring mbarrier words being initialised
    static_assert((2 * STAGES + 4 + 2 * NUM_EPI_WARPS) * 8 <= 496, "barrier area overflows into the TMEM slot");
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::BAR_OFFSET + 496);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("molly gemm: dynamic smem base not 1024-B aligned\n");
            __trap();
        }
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        if constexpr (EPI != EPI_SCATTER) tma_prefetch_desc(&tma_c);
        if constexpr (EPI == EPI_BIAS_RESID) tma_prefetch_desc(&tma_r);     // the residual source (== tma_c when in place)
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], CTA2 ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);   // pair: both CTAs' epilogues arrive
        }
        for (int s = 0; s < 2 * NUM_EPI_WARPS; ++s) mbar_init(&res_bar[s], 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        if constexpr (CTA2) { tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();                          // the TMEM base written by tcgen05.alloc is published inside the CTA by bar.sync
    if constexpr (CTA2) cluster_sync_all();   // (what compute-sanitizer racecheck models) ... and the barriers of BOTH CTAs are initialised
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int tiles_m = (p.M + TILE_M - 1) / TILE_M;
    const int splits = p.splits > 1 ? p.splits : 1;
    const int num_tiles = tiles_m * tiles_n * splits;
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int kb_per = (num_kb + splits - 1) / splits;        // the host picks `splits` so that no split is empty

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (GEMM_ELECT ? elect_one() : lane == 0) {   // (elect.sync: ptxas issues the TMA / MMA instructions of a one-thread region back to back)
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int ks = tile % splits, mn = tile / splits;
                const int m_blk = mn / tiles_n, n_blk = mn % tiles_n;
                const int a_row = m_blk * TILE_M + static_cast<int>(cta_rank) * BLOCK_M;
                const int b_row = n_blk * BLOCK_N + (CTA2 ? static_cast<int>(cta_rank) * (BLOCK_N / 2) : 0);
                const int kb_end = min(num_kb, (ks + 1) * kb_per);
                for (int kb = ks * kb_per; kb < kb_end; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* const a_dst = sA + stage * A_STAGE_BYTES;
                    uint8_t* const b_dst = sB + stage * Cfg::B_STAGE_BYTES;
                    if constexpr (CTA2) {
                        // both CTAs' bytes are credited to the leader's barrier; only the leader arms it
                        if (cta_rank == 0)
                            mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_STAGE_BYTES + Cfg::B_STAGE_BYTES));
                        if constexpr (A_MN) {
#pragma unroll
                            for (int bx = 0; bx < BLOCK_M / 64; ++bx)
                                tma_load_2d_pair(a_dst + bx * MN_BOX_BYTES, &tma_a, &full_bar[stage], a_row + bx * 64, kb * BLOCK_K);
                        } else {
                            tma_load_2d_pair(a_dst, &tma_a, &full_bar[stage], kb * BLOCK_K, a_row);
                        }
                        if constexpr (B_MN) {
#pragma unroll
                            for (int bx = 0; bx < BLOCK_N / 2 / 64; ++bx)
                                tma_load_2d_pair(b_dst + bx * MN_BOX_BYTES, &tma_b, &full_bar[stage], b_row + bx * 64, kb * BLOCK_K);
                        } else {
                            tma_load_2d_pair(b_dst, &tma_b, &full_bar[stage], kb * BLOCK_K, b_row);
                        }
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], A_STAGE_BYTES + Cfg::B_STAGE_BYTES);
                        if constexpr (A_MN) {
#pragma unroll
                            for (int bx = 0; bx < BLOCK_M / 64; ++bx)
                                tma_load_2d(a_dst + bx * MN_BOX_BYTES, &tma_a, &full_bar[stage], a_row + bx * 64, kb * BLOCK_K);
                        } else {
                            tma_load_2d(a_dst, &tma_a, &full_bar[stage], kb * BLOCK_K, a_row);
                        }
                        if constexpr (B_MN) {
#pragma unroll
                            for (int bx = 0; bx < BLOCK_N / 64; ++bx)
                                tma_load_2d(b_dst + bx * MN_BOX_BYTES, &tma_b, &full_bar[stage], b_row + bx * 64, kb * BLOCK_K);
                        } else {
#pragma unroll
                            for (int bx = 0; bx < BLOCK_N / B_BOX_ROWS; ++bx)      // the weight map has 128-row boxes
                                tma_load_2d(b_dst + bx * B_BOX_ROWS * BLOCK_K * 2, &tma_b, &full_bar[stage], kb * BLOCK_K,
                                            b_row + bx * B_BOX_ROWS);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (cta_rank == 0 && (GEMM_ELECT ? elect_one() : lane == 0)) {
            constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BLOCK_N, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++local) {
                const int acc = local & 1;
                const uint32_t acc_phase = (local >> 1) & 1;
                if constexpr (CTA2) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
                else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                const int kb_begin = (tile % splits) * kb_per, kb_end = min(num_kb, kb_begin + kb_per);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES), b_addr = smem_u32(sB + stage * Cfg::B_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // K-major: 16 K-elements = 32 B further inside the 128-B row; MN-major: 16 K-rows = 2 KB further down
                        const uint64_t a_desc = A_MN ? make_smem_desc(a_addr + k * UMMA_K * 128, MN_BOX_BYTES, 1024, kLayoutSW128)
                                                     : make_smem_desc(a_addr + k * UMMA_K * 2, 16, 8 * BLOCK_K * 2, kLayoutSW128);
                        const uint64_t b_desc = B_MN ? make_smem_desc(b_addr + k * UMMA_K * 128, MN_BOX_BYTES, 1024, kLayoutSW128)
                                                     : make_smem_desc(b_addr + k * UMMA_K * 2, 16, 8 * BLOCK_K * 2, kLayoutSW128);
                        const uint32_t accumulate = (kb > kb_begin || k != 0) ? 1u : 0u;
                        if constexpr (CTA2) umma_bf16_ss_pair(d_tmem, a_desc, b_desc, idesc, accumulate);
                        else umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, accumulate);
                    }
                    // frees this smem stage (in both CTAs of a pair) once the MMAs have read it
                    if constexpr (CTA2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue (of both CTAs of a pair)
                if constexpr (CTA2) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue ------------------------------
        const int ew = warp - 4;
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int half = ew >> 2;                   // which half of the tile's columns
        constexpr int COLS_PER_WARP = BLOCK_N / 2;
        uint8_t* const stg = smem + Cfg::STG_OFFSET + ew * (Cfg::STG_BUFS * STG_BUF_BYTES);
        uint64_t* const rbar = res_bar + 2 * ew;
        uint32_t rphase0 = 0, rphase1 = 0;
        const uint32_t lane_tmem = static_cast<uint32_t>(q * 32) << 16;
        int local = 0;
        for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++local) {
            const int m_blk = (tile / splits) / tiles_n, n_blk = (tile / splits) % tiles_n;
            const int acc = local & 1;
            const uint32_t acc_phase = (local >> 1) & 1;
            const int row0 = m_blk * TILE_M + static_cast<int>(cta_rank) * BLOCK_M + q * 32;   // first row of this warp's slab
            const int colw = n_blk * BLOCK_N + half * COLS_PER_WARP;     // first column of this warp's slab
            const uint32_t tacc = tmem_base + lane_tmem + acc * BLOCK_N + half * COLS_PER_WARP;

            if constexpr (EPI == EPI_SCATTER) {
                // ---- projector: rows go to hidden_states[b, start+1+j, :]; direct 16-B stores (0.3 % of the FLOPs)
                OutT* const out = reinterpret_cast<OutT*>(p.out);
                const int row = row0 + lane;
                long long dst_row = -1;
                if (row < p.M) {
                    const int n = row / p.seq_k;
                    const int j = row - n * p.seq_k;
                    const int b = __ldg(p.seq_table + 2 * n);
                    const int start = __ldg(p.seq_table + 2 * n + 1);
                    if (start >= 0 && j < p.k_cap) {
                        const int t = start + 1 + j;
                        if (t < p.T && b >= 0 && b < p.B) dst_row = static_cast<long long>(b) * p.T + t;
                        else if (p.err_flag) atomicOr(p.err_flag, 2);
                    }
                }
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                    const int col0 = colw + c * 32;
                    if (col0 >= p.N) break;                                   // warp-uniform
                    uint32_t raw[32];
                    tmem_ld32(tacc + c * 32, raw);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                    if (p.bias != nullptr) add_bias32(v, p.bias, col0);
                    if (dst_row >= 0) store_direct<OutT>(out + dst_row * p.ldo + col0, v);
                }
            } else if constexpr (EPI == EPI_BIAS_RESID) {
                // ---- out(fp32) = acc + bias + residual; residual chunks (32 rows x 32 fp32) arrive by TMA.
                // STG_BUFS == 2: chunk c+1 is prefetched while chunk c is processed (and chunk 0 under the main loop);
                // STG_BUFS == 1: one buffer, load -> add -> store per chunk (long-K GEMMs, where the ring needs the smem).
                constexpr int CHUNKS = COLS_PER_WARP / 32;
                int nchunks = 0;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) nchunks += (colw + c * 32 < p.N) ? 1 : 0;
                if (lane == 0 && nchunks > 0) {             // chunk 0 is fetched while the main loop still runs
                    tma_store_wait_read<0>();               // buffer 0 may still be read by an earlier store
                    mbar_arrive_expect_tx(&rbar[0], STG_BUF_BYTES);
                    tma_load_2d(stg, &tma_r, &rbar[0], colw, row0);
                }
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < nchunks; ++c) {
                    const int buf = (STG_BUFS == 2) ? (c & 1) : 0;
                    uint8_t* const sbuf = stg + buf * STG_BUF_BYTES;
                    if constexpr (STG_BUFS == 2) {
                        if (lane == 0 && c + 1 < nchunks) {  // prefetch the next residual chunk into the other buffer
                            tma_store_wait_read<0>();        // ... once the store that last used it has read it out
                            mbar_arrive_expect_tx(&rbar[buf ^ 1], STG_BUF_BYTES);
                            tma_load_2d(stg + (buf ^ 1) * STG_BUF_BYTES, &tma_r, &rbar[buf ^ 1], colw + (c + 1) * 32, row0);
                        }
                    } else {
                        if (lane == 0 && c > 0) {
                            tma_store_wait_read<0>();
                            mbar_arrive_expect_tx(&rbar[0], STG_BUF_BYTES);
                            tma_load_2d(stg, &tma_r, &rbar[0], colw + c * 32, row0);
                        }
                    }
                    const int col0 = colw + c * 32;
                    uint32_t raw[32];
                    tmem_ld32(tacc + c * 32, raw);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                    if (p.bias != nullptr) add_bias32(v, p.bias, col0);
                    if (buf == 0) { mbar_wait(&rbar[0], rphase0); rphase0 ^= 1; }
                    else          { mbar_wait(&rbar[1], rphase1); rphase1 ^= 1; }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        uint4* ptr = stg_chunk(sbuf, lane, i);
                        const uint4 r = *ptr;
                        uint4 o;
                        o.x = __float_as_uint(v[4 * i] + __uint_as_float(r.x));
                        o.y = __float_as_uint(v[4 * i + 1] + __uint_as_float(r.y));
                        o.z = __float_as_uint(v[4 * i + 2] + __uint_as_float(r.z));
                        o.w = __float_as_uint(v[4 * i + 3] + __uint_as_float(r.w));
                        *ptr = o;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tma_c, sbuf, col0, row0);
                        tma_store_commit();
                    }
                }
            } else {
                // ---- bf16 / fp32 outputs through one staging buffer per warp
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                constexpr int ACC_PER_CHUNK = (EPI == EPI_GLU) ? 128 : (kOutF32 ? 32 : 64);   // accumulator columns / store
                constexpr int CHUNKS = COLS_PER_WARP / ACC_PER_CHUNK;
                static_assert(CHUNKS >= 1, "GLU epilogue needs BLOCK_N == 256");
#pragma unroll 1
                for (int c = 0; c < CHUNKS; ++c) {
                    const int col0 = colw + c * ACC_PER_CHUNK;
                    if (col0 >= p.N) break;                                   // warp-uniform
                    if constexpr (EPI == EPI_GLU) {
                        // 128 accumulator columns (x1_0, x2_0, x1_1, ...) -> 64 outputs silu(x1) * x2 = one 128-B row
                        uint4 packed[8];
                        uint32_t* pw = reinterpret_cast<uint32_t*>(packed);
#pragma unroll
                        for (int hh = 0; hh < 4; ++hh) {
                            uint32_t raw[32];
                            tmem_ld32(tacc + c * 128 + hh * 32, raw);
                            tmem_ld_wait();
                            float v[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                            if (p.bias != nullptr) add_bias32(v, p.bias, col0 + hh * 32);   // add_bias_fnn variants: interleaved like W
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float g0 = silu(v[4 * i]) * v[4 * i + 1];
                                const float g1 = silu(v[4 * i + 2]) * v[4 * i + 3];
                                pw[hh * 8 + i] = pack_bf16x2(g0, g1);
                            }
                        }
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) *stg_chunk(stg, lane, i) = packed[i];
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tma_c, stg, col0 >> 1, row0);
                            tma_store_commit();
                        }
                    } else if constexpr (kOutF32) {
                        uint32_t raw[32];
                        tmem_ld32(tacc + c * 32, raw);
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                        if (p.bias != nullptr && tile % splits == 0) add_bias32(v, p.bias, col0);
                        if (EPI == EPI_BIAS && col0 < p.scale_cols) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] *= p.scale;
                        }
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *stg_chunk(stg, lane, i) = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]),
                                                                  __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (splits > 1 || p.reduce_out) tma_reduce_add_2d(&tma_c, stg, col0, row0);   // partial tile / accumulate
                            else tma_store_2d(&tma_c, stg, col0, row0);
                            tma_store_commit();
                        }
                    } else {
                        // bf16: 64 accumulator columns -> one 128-B row
                        float v[64];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int colh = col0 + hh * 32;
                            uint32_t raw[32];
                            tmem_ld32(tacc + c * 64 + hh * 32, raw);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[hh * 32 + i] = __uint_as_float(raw[i]);
                            if (colh < p.N) {                                   // N % 32 == 0: halves are all-in or all-out
                                if (p.bias != nullptr) add_bias32(v + hh * 32, p.bias, colh);
                                if ((EPI == EPI_BIAS || EPI == EPI_BIAS_ROPE) && colh < p.scale_cols) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) v[hh * 32 + i] *= p.scale;
                                }
                                if constexpr (EPI == EPI_BIAS_GELU) {
#pragma unroll
                                    for (int i = 0; i < 32; ++i) v[hh * 32 + i] = gelu_erf(v[hh * 32 + i]);
                                }
                            }
                        }
                        if constexpr (EPI == EPI_BIAS_ROPE) {
                            if (col0 < p.rope_cols) {                           // warp-uniform: q and k columns only
                                const int t = (row0 + lane) % p.seq_k;
                                if (p.rope_head_dim == 64) rope_chunk64<64>(v, t, p.rope_cos_t, p.rope_sin_t, p.rope_len);
                                else if (p.rope_head_dim == 32) rope_chunk64<32>(v, t, p.rope_cos_t, p.rope_sin_t, p.rope_len);
                                else rope_chunk64<16>(v, t, p.rope_cos_t, p.rope_sin_t, p.rope_len);
                            }
                        }
                        uint4 packed[8];
                        uint32_t* pw = reinterpret_cast<uint32_t*>(packed);
#pragma unroll
                        for (int i = 0; i < 32; ++i) pw[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) *stg_chunk(stg, lane, i) = packed[i];
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tma_c, stg, col0, row0);      // rows >= M / cols >= N are clipped by the TMA unit
                            tma_store_commit();
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA2) mbar_arrive_remote(&tmem_empty[acc], 0);      // the leader owns the MMA issue
                else mbar_arrive(&tmem_empty[acc]);
            }
        }
        if (lane == 0) tma_store_wait<0>();          // all of this warp's stores are complete before the CTA retires
    }

    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // nobody leaves while the peer still uses its smem / TMEM
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CTA2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BLOCK_N, int EPI, typename OutT, int STG_BUFS, bool CTA2, int OPND = GEMM_OPND_KK>
int launch_gemm_impl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p,
                     cudaStream_t stream, const CUtensorMap* tr = nullptr) {
    using Cfg = GemmCfg<BLOCK_N, STG_BUFS, CTA2>;
    auto kernel = gemm_tcgen05_kernel<BLOCK_N, EPI, OutT, STG_BUFS, CTA2, OPND>;
    static bool configured = false;
    if (!configured) {
        MOLLY_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    constexpr int TILE_M = CTA2 ? 2 * BLOCK_M : BLOCK_M;
    const int tiles = ((p.M + TILE_M - 1) / TILE_M) * ((p.N + BLOCK_N - 1) / BLOCK_N) * (p.splits > 1 ? p.splits : 1);
    const int slots = CTA2 ? device_sm_count() / 2 : device_sm_count();
    const int grid = (tiles < slots ? tiles : slots) * (CTA2 ? 2 : 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTA2 ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    {
        ProfScope prof(gemm_family(), 2.0 * p.M * p.N * p.K, stream);
        MOLLY_CUDA(cudaLaunchKernelEx(&cfg, kernel, ta, tb, tc, tr != nullptr ? *tr : tc, p));
    }
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

bool g_pair_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOLLY_GEMM_PAIR");          // bring-up switch: MOLLY_GEMM_PAIR=0 forces single-CTA tiles
        v = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}

template <int BLOCK_N, int EPI, typename OutT, int STG_BUFS = 1>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, bool pair,
                cudaStream_t stream, const CUtensorMap* tr = nullptr) {
    if constexpr (BLOCK_N == 256 && EPI != EPI_SCATTER) {
        if (pair) return launch_gemm_impl<256, EPI, OutT, STG_BUFS, true>(ta, tb, tc, p, stream, tr);
    }
    return launch_gemm_impl<BLOCK_N, EPI, OutT, STG_BUFS, false>(ta, tb, tc, p, stream, tr);
}

template <int BLOCK_N>
int dispatch_epilogue(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, int epi,
                      int out_dtype, bool pair, cudaStream_t stream, const CUtensorMap* tr) {
    const bool f32 = out_dtype == DT_F32;
    switch (epi) {
        case EPI_BIAS:
            return f32 ? launch_gemm<BLOCK_N, EPI_BIAS, float>(ta, tb, tc, p, pair, stream)
                       : launch_gemm<BLOCK_N, EPI_BIAS, __nv_bfloat16>(ta, tb, tc, p, pair, stream);
        case EPI_BIAS_GELU:
            return launch_gemm<BLOCK_N, EPI_BIAS_GELU, __nv_bfloat16>(ta, tb, tc, p, pair, stream);
        case EPI_BIAS_RESID:       // short K: the epilogue is a large share -> prefetch residual chunks; long K: deeper ring
            return p.K >= 2048 ? launch_gemm<BLOCK_N, EPI_BIAS_RESID, float, 1>(ta, tb, tc, p, pair, stream, tr)
                               : launch_gemm<BLOCK_N, EPI_BIAS_RESID, float, 2>(ta, tb, tc, p, pair, stream, tr);
        case EPI_BIAS_ROPE:
            return launch_gemm<BLOCK_N, EPI_BIAS_ROPE, __nv_bfloat16>(ta, tb, tc, p, pair, stream);
        case EPI_GLU:
            if constexpr (BLOCK_N == 256) {
                return launch_gemm<256, EPI_GLU, __nv_bfloat16>(ta, tb, tc, p, pair, stream);
            } else {
                MOLLY_CHECK(false, MOLLY_ERR_UNSUPPORTED, "gemm: GLU epilogue needs 256-wide tiles");
            }
        case EPI_SCATTER:
            return f32 ? launch_gemm<BLOCK_N, EPI_SCATTER, float>(ta, tb, tc, p, pair, stream)
                       : launch_gemm<BLOCK_N, EPI_SCATTER, __nv_bfloat16>(ta, tb, tc, p, pair, stream);
        default:
            MOLLY_CHECK(false, MOLLY_ERR_INVALID, "gemm: unknown epilogue %d", epi);
    }
}

}  // namespace

// Tile shape per problem: a CTA pair (cta_group::2) on 256 x 256, one CTA on 128 x 256 or one CTA on 128 x 128.  The pair
// is the fastest per FLOP (half the B traffic per SM), but small problems (generate-sized batches) fill the GPU better with
// smaller tiles.  Cost model: waves x per-SM work of one tile / measured relative efficiency of the shape.
GemmTile gemm_pick_tile(int M, int N, int epi) {
    const int sms = device_sm_count();
    auto cdiv = [](int a, int b) { return (a + b - 1) / b; };
    const bool pair_ok = g_pair_enabled() && epi != EPI_SCATTER && N % 256 == 0;
    const double c_pair = pair_ok ? cdiv(cdiv(M, 256) * cdiv(N, 256), sms / 2) * 2.0 : 1e30;
    const double c_256 = cdiv(cdiv(M, 128) * cdiv(N, 256), sms) * 2.0 / 0.90;
    const double c_128 = epi == EPI_GLU ? 1e30 : cdiv(cdiv(M, 128) * cdiv(N, 128), sms) * 1.0 / 0.70;
    if (c_pair <= c_256 && c_pair <= c_128) return GEMM_TILE_PAIR_256;
    return c_256 <= c_128 ? GEMM_TILE_256 : GEMM_TILE_128;
}

int gemm_make_map_a(CUtensorMap* ta, const void* a, int lda, int M, int K) {
    MOLLY_CHECK(K % 8 == 0 && lda % 8 == 0, MOLLY_ERR_UNSUPPORTED,
                "gemm: K and lda must be multiples of 8 (16-B TMA strides); got K=%d lda=%d", K, lda);
    MOLLY_CHECK((reinterpret_cast<uintptr_t>(a) & 15) == 0, MOLLY_ERR_INVALID, "gemm: A must be 16-B aligned");
    return make_tma_2d(ta, a, M, K, lda, BLOCK_M, BLOCK_K, 2);
}

int gemm_make_map_b(CUtensorMap* tb, const void* w, int ldw, int N, int K, int epi) {
    MOLLY_CHECK(K % 8 == 0 && ldw % 8 == 0, MOLLY_ERR_UNSUPPORTED,
                "gemm: K and ldw must be multiples of 8 (16-B TMA strides); got K=%d ldw=%d", K, ldw);
    MOLLY_CHECK((reinterpret_cast<uintptr_t>(w) & 15) == 0, MOLLY_ERR_INVALID, "gemm: W must be 16-B aligned");
    // 128-row boxes: a CTA pair loads one per CTA, a single-CTA tile one or two (rows past N are zero-filled by the TMA unit)
    (void)epi;
    return make_tma_2d(tb, w, N, K, ldw, B_BOX_ROWS, BLOCK_K, 2);
}

// Output (and residual) map: 32-row x 128-byte boxes, 128-B swizzle.  `n_out` = columns of the OUTPUT matrix.
int gemm_make_map_c(CUtensorMap* tc, void* out, int out_dtype, int ldo, int M, int n_out) {
    const int eb = out_dtype == DT_F32 ? 4 : 2;
    MOLLY_CHECK((static_cast<long long>(ldo) * eb) % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                MOLLY_ERR_INVALID, "gemm: output must be 16-B aligned with a 16-B multiple pitch (ldo=%d)", ldo);
    return make_tma_2d(tc, out, M, n_out, ldo, 32, 128 / eb, eb);
}

int gemm_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* tc, int M, int N, int K, int epi,
                const float* bias, void* out, int out_dtype, int ldo, const int32_t* seq_table, int seq_k, int B, int T,
                int k_cap, int32_t* err_flag, cudaStream_t stream, int scale_cols, float scale, const float* rope_cos_t,
                const float* rope_sin_t, int rope_len, int rope_cols, int rope_head_dim, const CUtensorMap* tr) {
    MOLLY_CHECK(M > 0 && N > 0 && K > 0, MOLLY_ERR_INVALID, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    if (epi == EPI_BIAS_ROPE) {
        MOLLY_CHECK(out_dtype == DT_BF16 && rope_cos_t != nullptr && rope_sin_t != nullptr && seq_k > 0 &&
                        rope_len >= seq_k, MOLLY_ERR_INVALID,
                    "gemm: rotary epilogue needs bf16 output, cos/sin tables covering k_tokens (%d >= %d)", rope_len, seq_k);
        MOLLY_CHECK((rope_head_dim == 16 || rope_head_dim == 32 || rope_head_dim == 64) && rope_cols % 64 == 0,
                    MOLLY_ERR_UNSUPPORTED, "gemm: rotary epilogue supports head_dim 16/32/64 (got %d), rope_cols %% 64 == 0",
                    rope_head_dim);
    }
    MOLLY_CHECK(N % 32 == 0, MOLLY_ERR_UNSUPPORTED, "gemm: N must be a multiple of 32, got %d", N);
    MOLLY_CHECK(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, MOLLY_ERR_INVALID,
                "gemm: bias must be 16-B aligned");
    MOLLY_CHECK(scale_cols % 32 == 0, MOLLY_ERR_UNSUPPORTED, "gemm: scale_cols must be a multiple of 32");
    if (epi == EPI_GLU) MOLLY_CHECK(N % 256 == 0, MOLLY_ERR_UNSUPPORTED, "gemm: GLU needs N %% 256 == 0 (N=%d)", N);
    if (epi == EPI_BIAS_GELU || epi == EPI_GLU)
        MOLLY_CHECK(out_dtype == DT_BF16, MOLLY_ERR_UNSUPPORTED, "gemm: GELU / GLU epilogues write bf16");
    if (epi == EPI_BIAS_RESID)
        MOLLY_CHECK(out_dtype == DT_F32, MOLLY_ERR_UNSUPPORTED, "gemm: the residual epilogue updates the fp32 stream");
    CUtensorMap dummy;
    if (epi == EPI_SCATTER) {
        MOLLY_CHECK(seq_table != nullptr && seq_k > 0 && B > 0 && T > 0 && out != nullptr, MOLLY_ERR_INVALID,
                    "gemm: scatter epilogue needs seq_table / k_tokens / B / T / out");
        MOLLY_CHECK(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, MOLLY_ERR_INVALID,
                    "gemm: scatter output must be 16-B aligned with ldo %% 8 == 0 (ldo=%d)", ldo);
        memset(&dummy, 0, sizeof(dummy));
        tc = &dummy;
    } else {
        MOLLY_CHECK(tc != nullptr, MOLLY_ERR_INVALID, "gemm: missing output tensor map");
    }
    GemmParams p{M, N, K, bias, out, ldo, seq_table, seq_k, B, T, k_cap, err_flag, scale_cols, scale,
                 rope_cos_t, rope_sin_t, rope_len, rope_cols, rope_head_dim};
    if (epi == EPI_BIAS_ACCUM) {
        MOLLY_CHECK(out_dtype == DT_F32, MOLLY_ERR_UNSUPPORTED, "gemm: the accumulating epilogue writes fp32");
        p.reduce_out = 1;
        epi = EPI_BIAS;
    }
    const GemmTile tile = gemm_pick_tile(M, N, epi);
    if (tile == GEMM_TILE_128) return dispatch_epilogue<128>(ta, tb, *tc, p, epi, out_dtype, false, stream, tr);
    return dispatch_epilogue<256>(ta, tb, *tc, p, epi, out_dtype, tile == GEMM_TILE_PAIR_256, stream, tr);
}

int gemm_make_map_mn(CUtensorMap* t, const void* base, int rows_k, int cols_mn, int ld) {
    MOLLY_CHECK(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, MOLLY_ERR_UNSUPPORTED,
                "gemm: an MN-major operand needs a 16-B aligned base and a pitch that is a multiple of 8 (ld=%d)", ld);
    return make_tma_2d(t, base, rows_k, cols_mn, ld, 64, 64, 2);
}

namespace {
// K splits of a wgrad: fewest waves per split (ceil(tiles * s / slots) / s) with a small charge per extra partial tile
int pick_splits(int tiles, int slots, int num_kb) {
    static const int max_splits = [] { const char* e = getenv("MOLLY_WGRAD_MAX_SPLITS"); return e ? atoi(e) : 8; }();
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= max_splits && s * 4 <= num_kb; ++s) {
        const int kb_per = (num_kb + s - 1) / s;
        if ((num_kb + kb_per - 1) / kb_per != s) continue;                 // an empty split would leave a tile unwritten
        const double cost = static_cast<double>((tiles * s + slots - 1) / slots) / s * (1.0 + 0.01 * (s - 1));
        if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    }
    return best;
}
}  // namespace

int gemm_launch_mn(int opnd, const void* a, int lda, const void* b, int ldb, int M, int N, int K, void* out, int out_dtype,
                   int ldo, cudaStream_t stream, bool out_zeroed) {
    MOLLY_CHECK(M > 0 && N > 0 && K > 0 && N % 32 == 0, MOLLY_ERR_UNSUPPORTED, "gemm_mn: M=%d N=%d K=%d (N %% 32 == 0)", M, N, K);
    MOLLY_CHECK((opnd == GEMM_OPND_K_MN && out_dtype == DT_BF16) || (opnd == GEMM_OPND_MN_MN && out_dtype == DT_F32),
                MOLLY_ERR_UNSUPPORTED, "gemm_mn: dgrad writes bf16, wgrad fp32 (opnd=%d dtype=%d)", opnd, out_dtype);
    CUtensorMap ta, tb, tc;
    int rc = opnd == GEMM_OPND_MN_MN ? gemm_make_map_mn(&ta, a, K, M, lda) : gemm_make_map_a(&ta, a, lda, M, K);
    if (rc) return rc;
    if ((rc = gemm_make_map_mn(&tb, b, K, N, ldb))) return rc;
    if ((rc = gemm_make_map_c(&tc, out, out_dtype, ldo, M, N))) return rc;
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.out = out; p.ldo = ldo; p.scale = 1.0f; p.splits = 1;
    const bool pair = g_pair_enabled() && N % 256 == 0 && M > BLOCK_M;
    if (opnd == GEMM_OPND_K_MN) {
        if (pair) return launch_gemm_impl<256, EPI_BIAS, __nv_bfloat16, 1, true, GEMM_OPND_K_MN>(ta, tb, tc, p, stream);
        return launch_gemm_impl<256, EPI_BIAS, __nv_bfloat16, 1, false, GEMM_OPND_K_MN>(ta, tb, tc, p, stream);
    }
    const int tile_m = pair ? 2 * BLOCK_M : BLOCK_M;
    const int tiles = ((M + tile_m - 1) / tile_m) * ((N + 255) / 256);
    p.splits = pick_splits(tiles, pair ? device_sm_count() / 2 : device_sm_count(), (K + BLOCK_K - 1) / BLOCK_K);
    if (p.splits > 1 && !out_zeroed) MOLLY_CUDA(cudaMemsetAsync(out, 0, static_cast<size_t>(M) * ldo * 4, stream));
    if (pair) return launch_gemm_impl<256, EPI_BIAS, float, 1, true, GEMM_OPND_MN_MN>(ta, tb, tc, p, stream);
    return launch_gemm_impl<256, EPI_BIAS, float, 1, false, GEMM_OPND_MN_MN>(ta, tb, tc, p, stream);
}

}  // namespace molly
