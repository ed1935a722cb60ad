// Attention backward (first piece of the encoder backward): d(q', k', v) from d(out) for the fused
// bidirectional MHA of attention.cu.  q' is the scaled + rotated query as stored in the packed QKV activation, so no score
// scaling appears here; the inverse rotation / scale belongs to the rotary backward.
//
// With P = softmax(S), S = q' k'^T (keys masked like the forward), dP = dO V^T, delta = rowsum(dO * O):
//     dS = P * (dP - delta)        dQ' = dS K'        dK' = dS^T Q'        dV = P^T dO
// Two kernels shaped like the forward one (TMA -> swizzled smem -> tcgen05.mma -> TMEM -> registers, thread = one row):
//   * attn_bwd_dq_kernel : CTA = 128 query rows; streams key blocks.  S and dP land in TMEM, the threads rebuild
//     P = exp2(S log2e - lse2) from the forward's row log-sum-exp (no running max, no rescaling), write dS as the bf16 K-major
//     A operand and dQ' += dS K' uses the K tile as the MN-major B operand (exactly how the forward feeds V).
//   * attn_bwd_dkv_kernel: CTA = 128 keys; streams query blocks with the roles swapped -- S^T = K' Q'^T, dP^T = V dO^T, so a
//     thread owns a KEY row, P^T and dS^T are K-major A operands, and dV += P^T dO, dK' += dS^T Q' take dO / Q' as MN-major B.
// The streamed blocks are SB = 64 rows for head_dim <= 64: the score tiles and accumulators then take 192 / 256 TMEM columns
// and ~80 / ~97 KB of shared memory, so TWO CTAs share an SM and one CTA's MMAs / barrier hand-overs run under the other's
// element-wise work (the first version streamed 128-row blocks with one CTA per SM: 5 500 cycles per 128 x 128 block against
// ~1 000 of MUFU work).  head_dim 128 keeps SB = 128, one CTA per SM (512 TMEM columns).  The element-wise work has no row
// reduction (P comes from the stored log-sum-exp), so TWO threads share a row, SB/2 columns each: warps w and w+4 address the
// same TMEM lanes.  Loads are double-buffered, the score tiles are not.
#include <math_constants.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

namespace {

constexpr int BW_BLOCK = 128;
constexpr int BW_THREADS = 288;                // 8 compute warps (two per query / key row: 64 columns each) + 1 control warp
constexpr int BW_COMPUTE = 256;
constexpr float BW_LOG2E = 1.4426950408889634f;
#ifndef BW_DS2
#define BW_DS2 1                               // 1: the dQ kernel double-buffers its dS tile (block j+1 is not held up by dQ MMA j)
#endif
#ifndef BW_TS
#define BW_TS 1                                // 1: the dK'/dV kernel keeps P^T / dS^T in tensor memory (A operand of the TS-form MMA)
#endif
#ifndef BW_ELECT
#define BW_ELECT 1                             // the control thread is chosen with elect.sync (see the note at the MMA loops)
#endif
#ifndef BW_WARP_ARRIVE
#define BW_WARP_ARRIVE 0                       // 1: one elected mbarrier.arrive per warp instead of one per thread
#endif
constexpr int BW_ARRIVALS = BW_WARP_ARRIVE ? BW_COMPUTE / 32 : BW_COMPUTE;

// every compute thread has done its part (TMEM reads fenced / smem writes published): signal the control thread
__device__ __forceinline__ void bw_arrive(uint64_t* bar) {
#if BW_WARP_ARRIVE
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
#else
    mbar_arrive(bar);
#endif
}

template <int D, int SB>
struct BwdCfg {
    static constexpr int BOX_D = D < 64 ? D : 64;
    static constexpr int NBOX = D / BOX_D;
    static constexpr int ROW_BYTES = BOX_D * 2;
    static constexpr uint32_t LAYOUT = ROW_BYTES == 128 ? kLayoutSW128 : (ROW_BYTES == 64 ? kLayoutSW64 : kLayoutSW32);
    static constexpr int BOX_BYTES = BW_BLOCK * ROW_BYTES;             // resident tiles: 128 rows per TMA box
    static constexpr int TILE = NBOX * BOX_BYTES;                      // one 128 x D bf16 tile
    static constexpr int SBOX_BYTES = SB * ROW_BYTES;                  // streamed tiles: SB rows per TMA box
    static constexpr int STILE = NBOX * SBOX_BYTES;                    // one SB x D bf16 tile
    static constexpr int PT_BYTES = BW_BLOCK * SB * 2;                 // one 128 x SB bf16 operand tile (SB/64 swizzle atoms)
    // streamed-tile ring depth.  Three stages when the blocks are small: a stage is refilled when the MMAs of the block BEFORE
    // the current one have completed, i.e. two iterations before its data is needed -- with two stages the TMA load was
    // issued one short iteration ahead and the control thread sat ~900 cycles on the full barrier (clock64 timeline)
    static constexpr int STAGES = D == 128 ? 1 : 3;                    // dK/dV kernel (shared memory: 2 x 112 KB per SM)
    static constexpr int DQ_STAGES = D == 128 ? 1 : 3;                 // dQ kernel (+ two dS tiles: 2 x 112 KB per SM)
    static constexpr int QPH = SB / 64;                                // 32-column quarters each of a row's two threads handles
    static constexpr int CTAS_PER_SM = SB == 64 ? 2 : 1;
    // dQ kernel: Q | dO | K ring | V ring | dS          TMEM: S | dP | dQ
    static constexpr int DQ_OFF_Q = 0, DQ_OFF_DO = TILE, DQ_OFF_K = 2 * TILE, DQ_OFF_V = DQ_OFF_K + DQ_STAGES * STILE;
    static constexpr int DS_BUFS = (BW_DS2 && DQ_STAGES >= 2) ? 2 : 1; // (the wait below pairs dS buffers with K/V stages)
    static constexpr int DQ_OFF_DS = DQ_OFF_V + DQ_STAGES * STILE, DQ_OFF_BAR = DQ_OFF_DS + DS_BUFS * PT_BYTES;
    static constexpr int DQ_SMEM = DQ_OFF_BAR + 256;
    static constexpr int DQ_TMEM = 2 * SB + D <= 256 ? 256 : 512;
    // dK/dV kernel: K | V | Q ring | dO ring | P^T | dS^T | row statistics of the streamed query block
    //                                                    TMEM: S^T | dP^T | dV | dK
    static constexpr int KV_OFF_K = 0, KV_OFF_V = TILE, KV_OFF_Q = 2 * TILE, KV_OFF_DO = KV_OFF_Q + STAGES * STILE;
    static constexpr int KV_OFF_PT = KV_OFF_DO + STAGES * STILE, KV_OFF_DST = KV_OFF_PT + PT_BYTES;
    static constexpr int KV_OFF_STAT = KV_OFF_DST + PT_BYTES, KV_OFF_BAR = KV_OFF_STAT + 2 * SB * 4;   // {lse2, delta} of SB queries
    static constexpr int KV_SMEM = KV_OFF_BAR + 256;
    static constexpr int KV_TMEM = 2 * SB + 2 * D <= 256 ? 256 : 512;
    static_assert(SB == 64 || SB == 128, "streamed blocks are 64 or 128 rows");
    static_assert(2 * SB + 2 * D <= 512, "tensor memory budget");
};

#ifdef BW_TIMELINE
// bring-up aid (-DBW_TIMELINE): clock64 stamps of the dQ kernel's compute thread 0 and control thread, first 16 CTAs
__device__ long long g_bw_tl[16 * 2 * 32 * 8];
#define BW_TL(role, it, slot)                                                                                         \
    do {                                                                                                              \
        const int _cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);                              \
        if (_cta < 16 && (it) < 32) g_bw_tl[((_cta * 2 + (role)) * 32 + (it)) * 8 + (slot)] = clock64();               \
    } while (0)
#else
#define BW_TL(role, it, slot) do {} while (0)
#endif

__device__ __forceinline__ float bw_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// delta[n, head, t] = sum_d dO[row, head*D + d] * O[row, head*D + d].  d/8 lanes share a (row, head) pair, 16 B each, so a warp
// reads 512 contiguous bytes of dO and of O per load instruction; the partial dot products meet in d/8-lane xor shuffles.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ d_out, const __nv_bfloat16* __restrict__ out, int n_seq, int k_tokens,
                  int heads, int d, float* __restrict__ delta) {
    const int lpp = d >> 3;                                             // lanes per (row, head) pair: 2, 4, 8 or 16
    const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long idx = gid / lpp, total = static_cast<long long>(n_seq) * k_tokens * heads;
    const int sub = static_cast<int>(gid % lpp);
    float acc = 0.f;
    if (idx < total) {                                                  // (pairs are contiguous: element offset = idx * d)
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(d_out + idx * d) + sub);
        const uint4 y = __ldg(reinterpret_cast<const uint4*>(out + idx * d) + sub);
        const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* yp = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 xf = __bfloat1622float2(xp[i]), yf = __bfloat1622float2(yp[i]);
            acc += xf.x * yf.x + xf.y * yf.y;
        }
    }
    for (int o = lpp >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (idx < total && sub == 0) {
        const int head = static_cast<int>(idx % heads);
        const long long row = idx / heads;
        const long long n = row / k_tokens, t = row % k_tokens;
        delta[(n * heads + head) * k_tokens + t] = acc;
    }
}

// store 32 bf16 values of operand-tile row r (K-major, 128-B swizzle, atom = 64 columns): columns [qt*32, qt*32 + 32)
__device__ __forceinline__ void store_quarter_row(uint8_t* tile, int r, int qt, const float* v) {
    uint8_t* row = tile + (qt >> 1) * (BW_BLOCK * 128) + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7, c0 = (qt & 1) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 u;
        u.x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]);
        u.y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
        u.z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]);
        u.w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
        *reinterpret_cast<uint4*>(row + (((c0 + c) ^ sw) << 4)) = u;
    }
}

// 128 x D fp32 accumulator rows from TMEM -> bf16 -> global (row-major, ld = ldo elements)
template <int D>
__device__ __forceinline__ void store_acc_rows(uint32_t tmem_acc, uint32_t lane_addr, __nv_bfloat16* dst, bool row_ok) {
#pragma unroll
    for (int c = 0; c < D / 16; ++c) {
        uint32_t o[16];
        tmem_ld16(tmem_acc + lane_addr + c * 16, o);
        tmem_ld_wait();
        if (row_ok) {
            uint4 u0, u1;
            u0.x = pack_bf16x2(__uint_as_float(o[0]), __uint_as_float(o[1]));
            u0.y = pack_bf16x2(__uint_as_float(o[2]), __uint_as_float(o[3]));
            u0.z = pack_bf16x2(__uint_as_float(o[4]), __uint_as_float(o[5]));
            u0.w = pack_bf16x2(__uint_as_float(o[6]), __uint_as_float(o[7]));
            u1.x = pack_bf16x2(__uint_as_float(o[8]), __uint_as_float(o[9]));
            u1.y = pack_bf16x2(__uint_as_float(o[10]), __uint_as_float(o[11]));
            u1.z = pack_bf16x2(__uint_as_float(o[12]), __uint_as_float(o[13]));
            u1.w = pack_bf16x2(__uint_as_float(o[14]), __uint_as_float(o[15]));
            reinterpret_cast<uint4*>(dst + c * 16)[0] = u0;
            reinterpret_cast<uint4*>(dst + c * 16)[1] = u1;
        }
    }
}

template <int D>
__device__ __forceinline__ void zero_row(__nv_bfloat16* dst) {
    for (int i = 0; i < D / 8; ++i) reinterpret_cast<uint4*>(dst)[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------- dQ'
// tma_qkv / tma_do: 128-row boxes (the CTA's own rows); tma_kv: SB-row boxes of the packed QKV activation (streamed K, V)
template <int D, int SB>
__global__ void __launch_bounds__(BW_THREADS, BwdCfg<D, SB>::CTAS_PER_SM)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                   const __grid_constant__ CUtensorMap tma_kv, int k_tokens,
                   int h, const int32_t* __restrict__ kv_info, const uint8_t* __restrict__ key_mask,
                   const float* __restrict__ lse2, const float* __restrict__ delta, __nv_bfloat16* __restrict__ d_qkv) {
    using Cfg = BwdCfg<D, SB>;
    constexpr int NST = Cfg::DQ_STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::DQ_OFF_BAR);
    uint64_t* bar_qdo = bars + 0;
    uint64_t* bar_kv_full = bars + 1;        // [NST <= 4]
    uint64_t* bar_kv_empty = bars + 5;       // [NST <= 4]  dQ MMA of the block done: the K/V stage and the dS tile are free
    uint64_t* bar_sdp_full = bars + 9;       // S and dP of the block are in TMEM
    uint64_t* bar_s_free = bars + 10;        // 256 arrivals: both are in registers
    // 256 arrivals: dS is in smem.  ONE barrier per dS tile: with two tiles a fast thread finishes block j+1 while a slow one is
    // still inside block j, and on a single barrier its second arrival would complete block j's phase in the slow thread's place
    // (found by running the parity tests under compute-sanitizer, whose instrumentation spreads the threads that far apart)
    uint64_t* bar_ds_full = bars + 11;       // [DS_BUFS <= 2]
    uint64_t* bar_dq_full = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BW_BLOCK, head = blockIdx.y, n = blockIdx.z, heads = gridDim.y;
    const int kvl = kv_info[2 * n], n_nonpad = kv_info[2 * n + 1];
    const bool interior = n_nonpad != kvl;
    const int nkv = (kvl + SB - 1) / SB;
    const long long row_base = static_cast<long long>(n) * k_tokens;
    if (nkv == 0) {                                          // no valid key: the forward wrote zeros, nothing flows back
        if (threadIdx.x < BW_BLOCK && q0 + threadIdx.x < k_tokens)
            zero_row<D>(d_qkv + (row_base + q0 + threadIdx.x) * 3 * h + head * D);
        return;
    }
    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tma_qkv);
            tma_prefetch_desc(&tma_do);
            mbar_init(bar_qdo, 1);
            for (int s = 0; s < NST; ++s) { mbar_init(&bar_kv_full[s], 1); mbar_init(&bar_kv_empty[s], 1); }
            mbar_init(bar_sdp_full, 1);
            mbar_init(bar_s_free, BW_ARRIVALS);
            for (int b = 0; b < Cfg::DS_BUFS; ++b) mbar_init(&bar_ds_full[b], BW_ARRIVALS);
            mbar_init(bar_dq_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, Cfg::DQ_TMEM);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + SB, tmem_dq = tmem_base + 2 * SB;
    // TS: dS (bf16, 32 packed columns per tile, two tiles) lives in the 64 tensor-memory columns the accumulators leave free and
    // feeds dQ += dS K as the A operand from tensor memory: no dS stores to shared memory, no proxy fence, a third of the
    // operand reads of that MMA.  Nothing is aliased, so S, dP of the next block are still issued under this block's work.
    constexpr bool TS = BW_TS && SB == 64 && Cfg::DS_BUFS == 2 && 2 * SB + D + 64 <= Cfg::DQ_TMEM;
    const uint32_t tmem_ds = tmem_base + 2 * SB + D;

    if (warp == 8) {
        if (BW_ELECT ? elect_one() : lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(BW_BLOCK, SB, false, false);
            constexpr uint32_t idesc_acc = make_idesc_bf16(BW_BLOCK, D, false, true);       // B MN-major
            const uint32_t s_q = smem_u32(smem + Cfg::DQ_OFF_Q), s_do = smem_u32(smem + Cfg::DQ_OFF_DO);
            const uint32_t s_k = smem_u32(smem + Cfg::DQ_OFF_K), s_v = smem_u32(smem + Cfg::DQ_OFF_V);
            const uint32_t s_ds = smem_u32(smem + Cfg::DQ_OFF_DS);
            auto load_tile = [&](int off, const CUtensorMap* map, uint64_t* bar, int col, int row) {
#pragma unroll
                for (int b = 0; b < Cfg::NBOX; ++b)
                    tma_load_2d(smem + off + b * Cfg::BOX_BYTES, map, bar, col + b * Cfg::BOX_D, row);
            };
            auto load_stile = [&](int off, uint64_t* bar, int col, int row) {
#pragma unroll
                for (int b = 0; b < Cfg::NBOX; ++b)
                    tma_load_2d(smem + off + b * Cfg::SBOX_BYTES, &tma_kv, bar, col + b * Cfg::BOX_D, row);
            };
            auto load_kv = [&](int blk) {
                const int st = blk % NST;
                mbar_arrive_expect_tx(&bar_kv_full[st], 2 * Cfg::STILE);
                load_stile(Cfg::DQ_OFF_K + st * Cfg::STILE, &bar_kv_full[st], h + head * D, static_cast<int>(row_base) + blk * SB);
                load_stile(Cfg::DQ_OFF_V + st * Cfg::STILE, &bar_kv_full[st], 2 * h + head * D, static_cast<int>(row_base) + blk * SB);
            };
            auto issue_scores = [&](int blk) {               // S = Q K^T and dP = dO V^T (all four operands K-major)
                const int st = blk % NST;
                mbar_wait(&bar_kv_full[st], (blk / NST) & 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < D / 16; ++s) {
                    const uint32_t off = ((s * 16) / Cfg::BOX_D) * Cfg::BOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    const uint32_t soff = ((s * 16) / Cfg::BOX_D) * Cfg::SBOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    umma_bf16_ss(tmem_s, make_smem_desc(s_q + off, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT),
                                 make_smem_desc(s_k + st * Cfg::STILE + soff, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT), idesc_s,
                                 s != 0);
                }
#pragma unroll
                for (int s = 0; s < D / 16; ++s) {
                    const uint32_t off = ((s * 16) / Cfg::BOX_D) * Cfg::BOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    const uint32_t soff = ((s * 16) / Cfg::BOX_D) * Cfg::SBOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    umma_bf16_ss(tmem_dp, make_smem_desc(s_do + off, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT),
                                 make_smem_desc(s_v + st * Cfg::STILE + soff, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT), idesc_s,
                                 s != 0);
                }
                umma_commit(bar_sdp_full);
            };
            mbar_arrive_expect_tx(bar_qdo, 2 * Cfg::TILE);
            load_tile(Cfg::DQ_OFF_Q, &tma_qkv, bar_qdo, head * D, static_cast<int>(row_base) + q0);
            load_tile(Cfg::DQ_OFF_DO, &tma_do, bar_qdo, head * D, static_cast<int>(row_base) + q0);
            for (int b = 0; b < NST && b < nkv; ++b) load_kv(b);
            mbar_wait(bar_qdo, 0);
            issue_scores(0);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % NST;
                BW_TL(1, j, 0);
                // refill the stage of block j-1 as soon as its dQ MMA (issued at the end of the previous iteration) is done:
                // the TMA load takes ~1 450 cycles to land and its block is needed NST-2 iterations from now
                if (NST >= 2 && j >= 1 && j - 1 + NST < nkv) {
                    mbar_wait(&bar_kv_empty[(j - 1) % NST], ((j - 1) / NST) & 1);
                    BW_TL(1, j, 5);
                    load_kv(j - 1 + NST);
                }
                if (NST >= 2 && j + 1 < nkv) {               // scores of block j+1 run under the element-wise work of block j
                    mbar_wait(bar_s_free, j & 1);
                    BW_TL(1, j, 1);
                    tc_fence_after();
                    issue_scores(j + 1);
                    BW_TL(1, j, 2);
                }
                mbar_wait(&bar_ds_full[j % Cfg::DS_BUFS], (j / Cfg::DS_BUFS) & 1);     // dS(j) is in smem
                BW_TL(1, j, 3);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < SB / 16; ++s) {          // dQ += dS K : A K-major (SB/64 atoms of 64 keys), B = K tile MN-major
                    const uint64_t bd = make_smem_desc(s_k + st * Cfg::STILE + s * 16 * Cfg::ROW_BYTES, Cfg::SBOX_BYTES,
                                                       8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
                    if constexpr (TS) {
                        umma_bf16_ts(tmem_dq, tmem_ds + (j % Cfg::DS_BUFS) * 32 + s * 8, bd, idesc_acc, (j | s) != 0);
                    } else {
                        const uint64_t ad = make_smem_desc(s_ds + (j % Cfg::DS_BUFS) * Cfg::PT_BYTES + (s >> 2) * (BW_BLOCK * 128) +
                                                               (s & 3) * 32, 16, 1024, kLayoutSW128);
                        umma_bf16_ss(tmem_dq, ad, bd, idesc_acc, (j | s) != 0);
                    }
                }
                umma_commit(&bar_kv_empty[st]);
                BW_TL(1, j, 4);
                if (j == nkv - 1) umma_commit(bar_dq_full);
                if (NST == 1 && j + 1 < nkv) {               // single stage: the next K/V overwrite this one after its dQ MMA
                    mbar_wait(&bar_kv_empty[0], j & 1);
                    load_kv(j + 1);
                    mbar_wait(bar_s_free, j & 1);
                    tc_fence_after();
                    issue_scores(j + 1);
                }
            }
        }
    } else {
        const int r = threadIdx.x & (BW_BLOCK - 1), half = threadIdx.x >> 7;     // row, and which 64 key columns of it
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const bool row_ok = q0 + r < k_tokens;
        const size_t stat = (static_cast<size_t>(n) * heads + head) * k_tokens + (row_ok ? q0 + r : 0);
        const float my_lse = row_ok ? lse2[stat] : CUDART_INF_F;             // +inf: P = 0 for rows outside the sequence
        const float my_delta = row_ok ? delta[stat] : 0.f;
        const float neg_lse = -my_lse;
        for (int j = 0; j < nkv; ++j) {
            if (threadIdx.x == 0) BW_TL(0, j, 0);
            mbar_wait(bar_sdp_full, j & 1);
            if (threadIdx.x == 0) BW_TL(0, j, 1);
            tc_fence_after();
            if (j >= Cfg::DS_BUFS) {                         // dQ MMA (j - DS_BUFS) has drained this dS tile
                const int jp = j - Cfg::DS_BUFS;
                mbar_wait(&bar_kv_empty[jp % NST], (jp / NST) & 1);
                tc_fence_after();
            }
            if (threadIdx.x == 0) BW_TL(0, j, 2);
            uint8_t* ds_tile = smem + Cfg::DQ_OFF_DS + (j % Cfg::DS_BUFS) * Cfg::PT_BYTES;
#pragma unroll
            for (int qq = 0; qq < Cfg::QPH; ++qq) {          // 32 keys at a time
                const int qt = Cfg::QPH * half + qq;
                uint32_t sr[32], pr[32];
                tmem_ld32(tmem_s + lane_addr + qt * 32, sr);
                tmem_ld32(tmem_dp + lane_addr + qt * 32, pr);
                tmem_ld_wait();
                if (threadIdx.x == 0) BW_TL(0, j, 3);
                if (qq == Cfg::QPH - 1) {
                    tc_fence_before();
                    bw_arrive(bar_s_free);                   // S / dP are in registers: the next block's scores may be issued
                }
                float v[32];
                if (!interior && (j + 1) * SB <= kvl) {      // every key of the block is valid (all but the last block): no masks
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float p = bw_ex2(fmaf(__uint_as_float(sr[i]), BW_LOG2E, neg_lse));
                        v[i] = p * (__uint_as_float(pr[i]) - my_delta);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int key = j * SB + qt * 32 + i;
                        bool ok = key < kvl;
                        if (interior && ok) ok = key_mask[row_base + key] != 0;
                        const float p = ok ? bw_ex2(fmaf(__uint_as_float(sr[i]), BW_LOG2E, neg_lse)) : 0.f;
                        v[i] = p * (__uint_as_float(pr[i]) - my_delta);
                    }
                }
                if (threadIdx.x == 0) BW_TL(0, j, 4);
                if constexpr (TS) {
                    uint32_t pk[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) pk[c] = pack_bf16x2(v[2 * c], v[2 * c + 1]);
                    tmem_st16(tmem_ds + lane_addr + (j % Cfg::DS_BUFS) * 32 + half * 16, pk);
                } else {
                    store_quarter_row(ds_tile, r, qt, v);
                }
            }
            if (threadIdx.x == 0) BW_TL(0, j, 5);
            if constexpr (TS) tmem_st_wait(); else fence_proxy_async_smem();
            tc_fence_before();
            bw_arrive(&bar_ds_full[j % Cfg::DS_BUFS]);
            if (threadIdx.x == 0) BW_TL(0, j, 6);
        }
        if (half == 0) {
            mbar_wait(bar_dq_full, 0);
            tc_fence_after();
            store_acc_rows<D>(tmem_dq, lane_addr, d_qkv + (row_base + q0 + r) * 3 * h + head * D, row_ok);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::DQ_TMEM);
    }
}

// ------------------------------------------------------------------------------------------------- dK', dV
// tma_qkv: 128-row boxes (the CTA's K, V); tma_q / tma_do: SB-row boxes (streamed Q', dO)
template <int D, int SB>
__global__ void __launch_bounds__(BW_THREADS, BwdCfg<D, SB>::CTAS_PER_SM)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_q,
                    const __grid_constant__ CUtensorMap tma_do, int k_tokens,
                    int h, const int32_t* __restrict__ kv_info, const uint8_t* __restrict__ key_mask,
                    const float* __restrict__ lse2, const float* __restrict__ delta, __nv_bfloat16* __restrict__ d_qkv) {
    using Cfg = BwdCfg<D, SB>;
    constexpr int NST = Cfg::STAGES;
    // TS: P^T and dS^T (bf16, 32 packed columns each) overwrite the first half of the fp32 S^T / dP^T columns they were computed
    // from and feed dV += P^T dO, dK' += dS^T Q' as the A operand FROM TENSOR MEMORY: no operand-tile stores, half the
    // shared-memory reads per accumulate MMA.  The next block's scores are then issued AFTER those MMAs (same thread: they
    // execute in order), so inside a CTA the block is a serial chain and the overlap comes from the second CTA of the SM.
    constexpr bool TS = BW_TS && SB == 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::KV_OFF_BAR);
    uint64_t* bar_kv = bars + 0;
    uint64_t* bar_q_full = bars + 1;         // [NST <= 3]  Q_i and dO_i landed
    uint64_t* bar_q_empty = bars + 4;        // [NST <= 3]  dV / dK MMAs of the block done: stage, P^T and dS^T tiles are free
    uint64_t* bar_sdp_full = bars + 7;
    uint64_t* bar_s_free = bars + 8;         // 256 arrivals
    uint64_t* bar_pt_full = bars + 9;        // 256 arrivals
    uint64_t* bar_out_full = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * BW_BLOCK, head = blockIdx.y, n = blockIdx.z, heads = gridDim.y;
    const int kvl = kv_info[2 * n], n_nonpad = kv_info[2 * n + 1];
    const bool interior = n_nonpad != kvl;
    const long long row_base = static_cast<long long>(n) * k_tokens;
    const int nq = (k_tokens + SB - 1) / SB;                 // every query row of the sequence attends (pad queries too)
    if (k0 >= kvl) {                                         // keys past the last valid one never receive probability mass
        if (threadIdx.x < BW_BLOCK && k0 + threadIdx.x < k_tokens) {
            __nv_bfloat16* row = d_qkv + (row_base + k0 + threadIdx.x) * 3 * h + head * D;
            zero_row<D>(row + h);
            zero_row<D>(row + 2 * h);
        }
        return;
    }
    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tma_qkv);
            tma_prefetch_desc(&tma_q);
            tma_prefetch_desc(&tma_do);
            mbar_init(bar_kv, 1);
            for (int s = 0; s < NST; ++s) { mbar_init(&bar_q_full[s], 1); mbar_init(&bar_q_empty[s], 1); }
            mbar_init(bar_sdp_full, 1);
            mbar_init(bar_s_free, BW_ARRIVALS);
            mbar_init(bar_pt_full, BW_ARRIVALS);
            mbar_init(bar_out_full, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, Cfg::KV_TMEM);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_st = tmem_base, tmem_dpt = tmem_base + SB, tmem_dv = tmem_base + 2 * SB, tmem_dk = tmem_dv + D;

    if (warp == 8) {
        if (BW_ELECT ? elect_one() : lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(BW_BLOCK, SB, false, false);
            constexpr uint32_t idesc_acc = make_idesc_bf16(BW_BLOCK, D, false, true);       // B MN-major
            const uint32_t s_k = smem_u32(smem + Cfg::KV_OFF_K), s_v = smem_u32(smem + Cfg::KV_OFF_V);
            const uint32_t s_q = smem_u32(smem + Cfg::KV_OFF_Q), s_do = smem_u32(smem + Cfg::KV_OFF_DO);
            const uint32_t s_pt = smem_u32(smem + Cfg::KV_OFF_PT), s_dst = smem_u32(smem + Cfg::KV_OFF_DST);
            auto load_tile = [&](int off, const CUtensorMap* map, uint64_t* bar, int col, int row) {
#pragma unroll
                for (int b = 0; b < Cfg::NBOX; ++b)
                    tma_load_2d(smem + off + b * Cfg::BOX_BYTES, map, bar, col + b * Cfg::BOX_D, row);
            };
            auto load_stile = [&](int off, const CUtensorMap* map, uint64_t* bar, int col, int row) {
#pragma unroll
                for (int b = 0; b < Cfg::NBOX; ++b)
                    tma_load_2d(smem + off + b * Cfg::SBOX_BYTES, map, bar, col + b * Cfg::BOX_D, row);
            };
            auto load_q = [&](int blk) {
                const int st = blk % NST;
                mbar_arrive_expect_tx(&bar_q_full[st], 2 * Cfg::STILE);
                load_stile(Cfg::KV_OFF_Q + st * Cfg::STILE, &tma_q, &bar_q_full[st], head * D, static_cast<int>(row_base) + blk * SB);
                load_stile(Cfg::KV_OFF_DO + st * Cfg::STILE, &tma_do, &bar_q_full[st], head * D, static_cast<int>(row_base) + blk * SB);
            };
            auto issue_scores = [&](int blk) {               // S^T = K Q^T and dP^T = V dO^T
                const int st = blk % NST;
                mbar_wait(&bar_q_full[st], (blk / NST) & 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < D / 16; ++s) {
                    const uint32_t off = ((s * 16) / Cfg::BOX_D) * Cfg::BOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    const uint32_t soff = ((s * 16) / Cfg::BOX_D) * Cfg::SBOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    umma_bf16_ss(tmem_st, make_smem_desc(s_k + off, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT),
                                 make_smem_desc(s_q + st * Cfg::STILE + soff, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT), idesc_s,
                                 s != 0);
                }
#pragma unroll
                for (int s = 0; s < D / 16; ++s) {
                    const uint32_t off = ((s * 16) / Cfg::BOX_D) * Cfg::BOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    const uint32_t soff = ((s * 16) / Cfg::BOX_D) * Cfg::SBOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    umma_bf16_ss(tmem_dpt, make_smem_desc(s_v + off, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT),
                                 make_smem_desc(s_do + st * Cfg::STILE + soff, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT), idesc_s,
                                 s != 0);
                }
                umma_commit(bar_sdp_full);
            };
            mbar_arrive_expect_tx(bar_kv, 2 * Cfg::TILE);
            load_tile(Cfg::KV_OFF_K, &tma_qkv, bar_kv, h + head * D, static_cast<int>(row_base) + k0);
            load_tile(Cfg::KV_OFF_V, &tma_qkv, bar_kv, 2 * h + head * D, static_cast<int>(row_base) + k0);
            for (int b = 0; b < NST && b < nq; ++b) load_q(b);
            mbar_wait(bar_kv, 0);
            issue_scores(0);
            for (int i = 0; i < nq; ++i) {
                const int st = i % NST;
                if (NST >= 2 && i >= 1 && i - 1 + NST < nq) {            // refill the stage of block i-1 (see the dQ kernel)
                    mbar_wait(&bar_q_empty[(i - 1) % NST], ((i - 1) / NST) & 1);
                    load_q(i - 1 + NST);
                }
                if (!TS && NST >= 2 && i + 1 < nq) {
                    mbar_wait(bar_s_free, i & 1);
                    tc_fence_after();
                    issue_scores(i + 1);
                }
                mbar_wait(bar_pt_full, i & 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < SB / 16; ++s) {          // contraction over the SB queries of the block
                    const uint64_t bo = make_smem_desc(s_do + st * Cfg::STILE + s * 16 * Cfg::ROW_BYTES, Cfg::SBOX_BYTES,
                                                       8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
                    const uint64_t bq = make_smem_desc(s_q + st * Cfg::STILE + s * 16 * Cfg::ROW_BYTES, Cfg::SBOX_BYTES,
                                                       8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
                    if constexpr (TS) {                      // A = 16 queries = 8 packed columns of tensor memory
                        umma_bf16_ts(tmem_dv, tmem_st + s * 8, bo, idesc_acc, (i | s) != 0);     // dV  += P^T  dO
                        umma_bf16_ts(tmem_dk, tmem_dpt + s * 8, bq, idesc_acc, (i | s) != 0);    // dK' += dS^T Q'
                    } else {
                        const uint32_t a_off = (s >> 2) * (BW_BLOCK * 128) + (s & 3) * 32;
                        const uint64_t pd = make_smem_desc(s_pt + a_off, 16, 1024, kLayoutSW128);
                        const uint64_t dd = make_smem_desc(s_dst + a_off, 16, 1024, kLayoutSW128);
                        umma_bf16_ss(tmem_dv, pd, bo, idesc_acc, (i | s) != 0);
                        umma_bf16_ss(tmem_dk, dd, bq, idesc_acc, (i | s) != 0);
                    }
                }
                umma_commit(&bar_q_empty[st]);
                if (i == nq - 1) umma_commit(bar_out_full);
                if (TS && i + 1 < nq) issue_scores(i + 1);   // behind the MMAs that read the columns S^T / dP^T(i+1) overwrite
                if (NST == 1 && i + 1 < nq) {
                    mbar_wait(&bar_q_empty[0], i & 1);
                    load_q(i + 1);
                    mbar_wait(bar_s_free, i & 1);
                    tc_fence_after();
                    issue_scores(i + 1);
                }
            }
        }
    } else {
        const int r = threadIdx.x & (BW_BLOCK - 1), half = threadIdx.x >> 7;     // key row k0 + r, 64 of the 128 queries
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int key = k0 + r;
        bool key_ok = key < kvl;
        if (interior && key_ok) key_ok = key_mask[row_base + key] != 0;
        uint8_t* pt_tile = smem + Cfg::KV_OFF_PT;
        uint8_t* dst_tile = smem + Cfg::KV_OFF_DST;
        const size_t stat_base = (static_cast<size_t>(n) * heads + head) * k_tokens;
        float* st_lse = reinterpret_cast<float*>(smem + Cfg::KV_OFF_STAT);
        float* st_delta = st_lse + SB;
        for (int i = 0; i < nq; ++i) {
            // row statistics of the SB queries of this block: thread r brings query i*SB + r (the load is in flight during the
            // waits below)
            const int q = i * SB + r;
            float my_stat = 0.f;
            if (r < SB) {
                if (half == 0) my_stat = q < k_tokens ? lse2[stat_base + q] : CUDART_INF_F;      // +inf: P = 0 beyond the sequence
                else my_stat = q < k_tokens ? delta[stat_base + q] : 0.f;
            }
            mbar_wait(bar_sdp_full, i & 1);
            tc_fence_after();
            if (!TS && i > 0) {                              // dV / dK MMAs (i-1) have drained the P^T / dS^T tiles ...
                mbar_wait(&bar_q_empty[(i - 1) % NST], ((i - 1) / NST) & 1);
                tc_fence_after();
            }                                                // (TS: sdp_full(i) was committed behind those MMAs by the same thread)
            // ... and, since they only ran after all 256 arrivals on pt_full(i-1), every thread is done reading the previous
            // block's statistics: the single buffer can be overwritten
            if (r < SB) (half == 0 ? st_lse : st_delta)[r] = my_stat;
            named_bar_sync(1, BW_COMPUTE);
#pragma unroll
            for (int qq = 0; qq < Cfg::QPH; ++qq) {          // 32 queries at a time
                const int qt = Cfg::QPH * half + qq;
                uint32_t sr[32], pr[32];
                tmem_ld32(tmem_st + lane_addr + qt * 32, sr);
                tmem_ld32(tmem_dpt + lane_addr + qt * 32, pr);
                tmem_ld_wait();
                if (!TS && qq == Cfg::QPH - 1) {
                    tc_fence_before();
                    bw_arrive(bar_s_free);
                }
                float pv[32], dv[32];
                if (!key_ok) {                               // a key past the sequence / masked: its whole row of P^T is zero
#pragma unroll
                    for (int c = 0; c < 32; ++c) sr[c] = 0xff800000u;             // S = -inf -> P = exp2(-inf) = 0
                }
                const float4* l4 = reinterpret_cast<const float4*>(st_lse + qt * 32);          // (broadcast reads)
                const float4* d4 = reinterpret_cast<const float4*>(st_delta + qt * 32);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 l = l4[c4], dl = d4[c4];
                    const float ls[4] = {l.x, l.y, l.z, l.w}, ds[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = c4 * 4 + e;
                        const float p = bw_ex2(fmaf(__uint_as_float(sr[c]), BW_LOG2E, -ls[e]));
                        pv[c] = p;
                        dv[c] = p * (__uint_as_float(pr[c]) - ds[e]);
                    }
                }
                if constexpr (TS) {
                    uint32_t pp[16], dp[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        pp[c] = pack_bf16x2(pv[2 * c], pv[2 * c + 1]);
                        dp[c] = pack_bf16x2(dv[2 * c], dv[2 * c + 1]);
                    }
                    // warps w and w+4 share these TMEM lanes: the other half's thread must have READ columns 16..31 of S^T / dP^T
                    // before this one's packed words land there
                    named_bar_sync(2, BW_COMPUTE);
                    tmem_st16(tmem_st + lane_addr + half * 16, pp);
                    tmem_st16(tmem_dpt + lane_addr + half * 16, dp);
                } else {
                    store_quarter_row(pt_tile, r, qt, pv);
                    store_quarter_row(dst_tile, r, qt, dv);
                }
            }
            if constexpr (TS) tmem_st_wait(); else fence_proxy_async_smem();
            tc_fence_before();
            bw_arrive(bar_pt_full);
        }
        mbar_wait(bar_out_full, 0);
        tc_fence_after();
        __nv_bfloat16* row = d_qkv + (row_base + key) * 3 * h + head * D;
        if (half == 0) store_acc_rows<D>(tmem_dk, lane_addr, row + h, key < k_tokens);
        else store_acc_rows<D>(tmem_dv, lane_addr, row + 2 * h, key < k_tokens);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::KV_TMEM);
    }
}

template <int D>
int launch_attention_bwd(const void* qkv, const void* d_out, int n_seq, int k_tokens, int h, int heads,
                         const int32_t* kv_info, const uint8_t* key_mask, const float* lse2, const float* delta,
                         __nv_bfloat16* d_qkv, cudaStream_t stream) {
    constexpr int SB = D <= 64 ? 64 : 128;
    using Cfg = BwdCfg<D, SB>;
    static bool configured = false;
    if (!configured) {
        MOLLY_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<D, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::DQ_SMEM));
        MOLLY_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<D, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::KV_SMEM));
        // two CTAs per SM need the whole shared-memory carve-out
        MOLLY_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<D, SB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
        MOLLY_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<D, SB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    const int rows = n_seq * k_tokens;
    CUtensorMap tq, tdo, tq_s, tdo_s;          // 128-row boxes for the CTA's own tile, SB-row boxes for the streamed ones
    int rc = make_tma_2d(&tq, qkv, rows, 3 * h, 3 * h, BW_BLOCK, Cfg::BOX_D, 2);
    if (rc) return rc;
    if ((rc = make_tma_2d(&tdo, d_out, rows, h, h, BW_BLOCK, Cfg::BOX_D, 2))) return rc;
    if ((rc = make_tma_2d(&tq_s, qkv, rows, 3 * h, 3 * h, SB, Cfg::BOX_D, 2))) return rc;
    if ((rc = make_tma_2d(&tdo_s, d_out, rows, h, h, SB, Cfg::BOX_D, 2))) return rc;
    dim3 grid((k_tokens + BW_BLOCK - 1) / BW_BLOCK, heads, n_seq);
    {   // 4 MMAs of 2*K*K*d each per (sequence, head) in each kernel... dense-equivalent 2.5x the forward in total
        prof_attention_work(kv_info, n_seq, k_tokens, h, 10.0, stream);     // work = 10 h K sum(kv_len), known on the device only
        ProfScope prof(PF_ATTENTION_BWD, 10.0 * n_seq * k_tokens * static_cast<double>(k_tokens) * h, stream);
        attn_bwd_dq_kernel<D, SB><<<grid, BW_THREADS, Cfg::DQ_SMEM, stream>>>(tq, tdo, tq_s, k_tokens, h, kv_info, key_mask, lse2,
                                                                              delta, d_qkv);
        attn_bwd_dkv_kernel<D, SB><<<grid, BW_THREADS, Cfg::KV_SMEM, stream>>>(tq, tq_s, tdo_s, k_tokens, h, kv_info, key_mask,
                                                                               lse2, delta, d_qkv);
    }
    count_launch();
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

}  // namespace

int attention_bwd_launch(const void* qkv, const void* out, const void* d_out, const float* lse2, int n_seq, int k_tokens, int h,
                         int heads, const int32_t* kv_info, const uint8_t* key_mask, void* d_qkv, float* delta_ws,
                         cudaStream_t stream) {
    MOLLY_CHECK(n_seq > 0 && k_tokens > 0 && heads > 0 && h % heads == 0, MOLLY_ERR_INVALID,
                "attention_bwd: bad shape n_seq=%d k=%d h=%d heads=%d", n_seq, k_tokens, h, heads);
    const int d = h / heads;
    MOLLY_CHECK(d == 16 || d == 32 || d == 64 || d == 128, MOLLY_ERR_UNSUPPORTED, "attention_bwd: head_dim %d unsupported", d);
    MOLLY_CHECK(static_cast<long long>(n_seq) * k_tokens < (1ll << 31) && n_seq <= 65535 && heads <= 65535,
                MOLLY_ERR_UNSUPPORTED, "attention_bwd: problem too large for the grid / TMA coordinates");
    const long long items = static_cast<long long>(n_seq) * k_tokens * heads * (d / 8);       // one thread per 16 B
    attn_delta_kernel<<<static_cast<unsigned>((items + 255) / 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(d_out), static_cast<const __nv_bfloat16*>(out), n_seq, k_tokens, heads, d, delta_ws);
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(d_qkv);
    switch (d) {
        case 16: return launch_attention_bwd<16>(qkv, d_out, n_seq, k_tokens, h, heads, kv_info, key_mask, lse2, delta_ws, dq, stream);
        case 32: return launch_attention_bwd<32>(qkv, d_out, n_seq, k_tokens, h, heads, kv_info, key_mask, lse2, delta_ws, dq, stream);
        case 64: return launch_attention_bwd<64>(qkv, d_out, n_seq, k_tokens, h, heads, kv_info, key_mask, lse2, delta_ws, dq, stream);
        default: return launch_attention_bwd<128>(qkv, d_out, n_seq, k_tokens, h, heads, kv_info, key_mask, lse2, delta_ws, dq, stream);
    }
}

}  // namespace molly

#ifdef BW_TIMELINE
extern "C" __attribute__((visibility("default"))) int molly_debug_bw_timeline(long long* host_out) {
    return static_cast<int>(cudaMemcpyFromSymbol(host_out, molly::g_bw_tl, sizeof(molly::g_bw_tl)));
}
#endif
