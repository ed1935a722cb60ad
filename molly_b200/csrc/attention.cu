// Fused bidirectional multi-head attention over K-token padded sequences with per-sequence key masking
// (HF:257-282 eager attention with the additive key-padding mask of HF:679-709), flash-style online softmax.
//
// Input is the packed bf16 QKV activation [n_seq*k_tokens, 3h] written by the fused QKV GEMM (q already scaled by
// d^-1/2 through the packed weights, rotary already applied); output is bf16 [n_seq*k_tokens, h].
//
// One CTA = 128 query rows of one (sequence, head):
//   * TMA brings Q once and K/V tiles (128 keys) through a 2-stage mbarrier ring
//   * S = Q K^T  : tcgen05.mma, both operands K-major from swizzled smem, fp32 accumulator in TMEM (128 columns)
//   * softmax    : 4 warps, thread r owns row r (tcgen05.ld 32x32b: TMEM lane == row, so row max / row sum need no
//                  shuffles); exp2 with the running max in the log2 domain; P written to smem as the bf16 K-major
//                  A operand (128-B swizzle); O rescaled in TMEM only when a warp's running max moved
//   * O += P V   : tcgen05.mma, B operand = V tile straight from TMA ([key][d] row-major == MN-major B), fp32 in TMEM
//   * keys >= kv_len are never loaded (whole blocks skipped); interior pad keys use the byte mask
// Two CTAs are co-resident per SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
#include <math_constants.h>
#include <stdlib.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

namespace {

constexpr int ATT_BLOCK = 128;                 // query rows per CTA == keys per KV block
constexpr int ATT_THREADS = 160;               // 4 softmax warps + 1 control warp
constexpr float LOG2E = 1.4426950408889634f;
constexpr int ATTN_POLY_DEFAULT = 0;           // set from the A/B measurement (tools/attn_bench.py)

template <int D>
struct AttnCfg {
    static constexpr int BOX_D = D < 64 ? D : 64;
    static constexpr int NBOX = D / BOX_D;
    static constexpr int ROW_BYTES = BOX_D * 2;                       // 32 / 64 / 128
    static constexpr uint32_t LAYOUT = ROW_BYTES == 128 ? kLayoutSW128 : (ROW_BYTES == 64 ? kLayoutSW64 : kLayoutSW32);
    static constexpr int BOX_BYTES = ATT_BLOCK * ROW_BYTES;
    static constexpr int TILE_BYTES = NBOX * BOX_BYTES;               // one Q / K / V tile
    static constexpr int P_BYTES = ATT_BLOCK * ATT_BLOCK * 2;         // two 128x64 bf16 swizzle atoms
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = TILE_BYTES;
    static constexpr int OFF_V = 3 * TILE_BYTES;
    static constexpr int OFF_P = 5 * TILE_BYTES;
    static constexpr int OFF_BAR = OFF_P + P_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 128;
    static constexpr int TMEM_COLS = 256;                             // S: 128, O: D (<= 128)
    static constexpr int MIN_CTAS = SMEM_BYTES <= 115000 ? 2 : 1;
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// exp2 on the FMA pipe (FlashAttention-4 style) for POLY out of every 4 element pairs: the softmax is bound by the
// 16-per-clock MUFU.EX2 unit, while the FMA pipe is ~80 % idle.  2^x = 2^n * p(r), n = round(x), r = x - n in [-0.5, 0.5],
// p = cubic with max relative error 1.0e-4 (P is rounded to bf16 right after: 3.9e-3), 2^n spliced into the exponent field.
__device__ __forceinline__ void exp2_poly_pair(float& x0, float& x1) {
    const uint64_t magic = pack_f32x2(12582912.0f, 12582912.0f);            // 1.5 * 2^23: low mantissa bits = round(x)
    const uint64_t x2 = pack_f32x2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));     // keep 2^n a normal number
    const uint64_t t2 = add_f32x2(x2, magic);
    const uint64_t n2 = add_f32x2(t2, pack_f32x2(-12582912.0f, -12582912.0f));
    const uint64_t r2 = fma_f32x2(n2, pack_f32x2(-1.0f, -1.0f), x2);
    uint64_t p2 = fma_f32x2(pack_f32x2(0.05583828315138817f, 0.05583828315138817f), r2,
                            pack_f32x2(0.2426394820213318f, 0.2426394820213318f));
    p2 = fma_f32x2(p2, r2, pack_f32x2(0.6931367516517639f, 0.6931367516517639f));
    p2 = fma_f32x2(p2, r2, pack_f32x2(0.9999245405197144f, 0.9999245405197144f));
    float p0, p1, t0, t1;
    unpack_f32x2(p2, p0, p1);
    unpack_f32x2(t2, t0, t1);
    x0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    x1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

template <int D, int POLY>
__global__ void __launch_bounds__(ATT_THREADS, AttnCfg<D>::MIN_CTAS)
attention_kernel(const __grid_constant__ CUtensorMap tma_qkv, int k_tokens, int h, const int32_t* __restrict__ kv_info,
                 const uint8_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out) {
    using Cfg = AttnCfg<D>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* bar_q = bars + 0;
    uint64_t* bar_kv_full = bars + 1;     // [2]
    uint64_t* bar_kv_empty = bars + 3;    // [2]
    uint64_t* bar_s_full = bars + 5;
    uint64_t* bar_p_full = bars + 6;
    uint64_t* bar_o_full = bars + 7;
    uint64_t* bar_s_free = bars + 8;      // softmax has S(j) in registers -> S(j+1) may overwrite the TMEM columns
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_BLOCK, head = blockIdx.y, n = blockIdx.z;
    const int kvl = kv_info[2 * n], n_nonpad = kv_info[2 * n + 1];
    const bool interior = n_nonpad != kvl;                  // pad ids before the last real token
    const int nkv = (kvl + ATT_BLOCK - 1) / ATT_BLOCK;
    const long long row_base = static_cast<long long>(n) * k_tokens;

    if (nkv == 0) {                                         // all-pad sequence: the reference never encodes one
        if (warp < 4 && q0 + threadIdx.x < k_tokens) {
            uint4* o = reinterpret_cast<uint4*>(out + (row_base + q0 + threadIdx.x) * h + head * D);
            for (int i = 0; i < D / 8; ++i) o[i] = make_uint4(0, 0, 0, 0);
        }
        return;
    }

    if (warp == 4) {
        if (lane == 0) {
            if ((smem_u32(smem) & 1023u) != 0) { printf("molly attention: smem base not 1024-B aligned\n"); __trap(); }
            tma_prefetch_desc(&tma_qkv);
            mbar_init(bar_q, 1);
            mbar_init(&bar_kv_full[0], 1); mbar_init(&bar_kv_full[1], 1);
            mbar_init(&bar_kv_empty[0], 1); mbar_init(&bar_kv_empty[1], 1);
            mbar_init(bar_s_full, 1);
            mbar_init(bar_p_full, 128);
            mbar_init(bar_o_full, 1);
            mbar_init(bar_s_free, 128);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + 128;

    if (warp == 4) {
        if (lane == 0) {
            // ---------------- control thread: TMA producer + MMA issuer ----------------
            auto load_tile = [&](int smem_off, uint64_t* bar, int col, int row) {
#pragma unroll
                for (int b = 0; b < Cfg::NBOX; ++b)
                    tma_load_2d(smem + smem_off + b * Cfg::BOX_BYTES, &tma_qkv, bar, col + b * Cfg::BOX_D, row);
            };
            const int qcol = head * D, kcol = h + head * D, vcol = 2 * h + head * D;
            constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BLOCK, ATT_BLOCK, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BLOCK, D, false, true);      // B (= V) is MN-major
            const uint32_t s_q = smem_u32(smem + Cfg::OFF_Q), s_k = smem_u32(smem + Cfg::OFF_K);
            const uint32_t s_v = smem_u32(smem + Cfg::OFF_V), s_p = smem_u32(smem + Cfg::OFF_P);
            auto load_kv = [&](int blk) {
                const int stg = blk & 1;
                mbar_arrive_expect_tx(&bar_kv_full[stg], 2 * Cfg::TILE_BYTES);
                load_tile(Cfg::OFF_K + stg * Cfg::TILE_BYTES, &bar_kv_full[stg], kcol,
                          static_cast<int>(row_base) + blk * ATT_BLOCK);
                load_tile(Cfg::OFF_V + stg * Cfg::TILE_BYTES, &bar_kv_full[stg], vcol,
                          static_cast<int>(row_base) + blk * ATT_BLOCK);
            };
            auto issue_s = [&](int blk) {        // S = Q K(blk)^T : K-major x K-major, D/16 k-steps
                mbar_wait(&bar_kv_full[blk & 1], (blk >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < D / 16; ++s) {
                    const uint32_t off = ((s * 16) / Cfg::BOX_D) * Cfg::BOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    const uint64_t qd = make_smem_desc(s_q + off, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
                    const uint64_t kd =
                        make_smem_desc(s_k + (blk & 1) * Cfg::TILE_BYTES + off, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
                    umma_bf16_ss(tmem_s, qd, kd, idesc_s, s != 0);
                }
                umma_commit(bar_s_full);
            };
            mbar_arrive_expect_tx(bar_q, Cfg::TILE_BYTES);
            load_tile(Cfg::OFF_Q, bar_q, qcol, static_cast<int>(row_base) + q0);
            load_kv(0);
            if (nkv > 1) load_kv(1);
            mbar_wait(bar_q, 0);
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                const int st = j & 1;
                // (1) the moment the softmax warps hold S(j) in registers, S(j+1) is issued: it runs under softmax(j)
                mbar_wait(bar_s_free, j & 1);
                tc_fence_after();
                if (j + 1 < nkv) issue_s(j + 1);
                // (2) O += P(j) V(j) : P K-major (two 64-key atoms), V MN-major; 8 k-steps of 16 keys
                mbar_wait(bar_p_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < ATT_BLOCK / 16; ++s) {
                    const uint64_t pd = make_smem_desc(s_p + (s >> 2) * (ATT_BLOCK * 128) + (s & 3) * 32, 16, 1024,
                                                       kLayoutSW128);
                    const uint64_t vd = make_smem_desc(s_v + st * Cfg::TILE_BYTES + s * 16 * Cfg::ROW_BYTES,
                                                       Cfg::BOX_BYTES, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
                    umma_bf16_ss(tmem_o, pd, vd, idesc_pv, (j | s) != 0);
                }
                umma_commit(&bar_kv_empty[st]);              // PV(j) done: stage st, the P buffer and O(j) are free
                if (j == nkv - 1) umma_commit(bar_o_full);
                // (3) refill stage st with block j+2 once PV(j) has drained it
                if (j + 2 < nkv) {
                    mbar_wait(&bar_kv_empty[st], (j >> 1) & 1);
                    load_kv(j + 2);
                }
            }
        }
    } else {
        // ---------------- softmax warps: thread r owns query row r ----------------
        const int r = threadIdx.x;
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        uint8_t* p_row = smem + Cfg::OFF_P + (r >> 3) * 1024 + (r & 7) * 128;
        const int sw = r & 7;
        float m_run = -CUDART_INF_F;      // running max, log2 domain
        float l_run = 0.f;
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(bar_s_full, j & 1);
            tc_fence_after();
            float s[ATT_BLOCK];
            {
                uint32_t raw[ATT_BLOCK];
                tmem_ld32(tmem_s + lane_addr + 0, raw);
                tmem_ld32(tmem_s + lane_addr + 32, raw + 32);
                tmem_ld32(tmem_s + lane_addr + 64, raw + 64);
                tmem_ld32(tmem_s + lane_addr + 96, raw + 96);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < ATT_BLOCK; ++i) s[i] = __uint_as_float(raw[i]);
            }
            tc_fence_before();
            mbar_arrive(bar_s_free);                         // S(j) is in registers: the MMA warp may start S(j+1)
            const int j0 = j * ATT_BLOCK;
            if (interior) {
                const uint4* mk = reinterpret_cast<const uint4*>(key_mask + row_base + j0);
                const bool vec_ok = ((row_base + j0) & 15) == 0 && j0 + ATT_BLOCK <= k_tokens;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint32_t w[4];
                    if (vec_ok) {
                        const uint4 u = __ldg(mk + g);
                        w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            w[i] = 0;
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const int c = j0 + g * 16 + i * 4 + b;
                                const uint32_t v = (c < k_tokens) ? key_mask[row_base + c] : 0;
                                w[i] |= (v & 0xffu) << (8 * b);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const bool ok = ((w[i >> 2] >> (8 * (i & 3))) & 0xffu) != 0 && (j0 + g * 16 + i < kvl);
                        if (!ok) s[g * 16 + i] = -CUDART_INF_F;
                    }
                }
            } else if (j0 + ATT_BLOCK > kvl) {
                const int lim = kvl - j0;
#pragma unroll
                for (int i = 0; i < ATT_BLOCK; ++i)
                    if (i >= lim) s[i] = -CUDART_INF_F;
            }
            float mx4[4] = {s[0], s[1], s[2], s[3]};          // 4 independent chains instead of one 127-deep one
#pragma unroll
            for (int i = 4; i < ATT_BLOCK; i += 4) {
                mx4[0] = fmaxf(mx4[0], s[i]); mx4[1] = fmaxf(mx4[1], s[i + 1]);
                mx4[2] = fmaxf(mx4[2], s[i + 2]); mx4[3] = fmaxf(mx4[3], s[i + 3]);
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // Lazy rescaling: the reference max only moves when the row max grew by more than 2^8; until then P is
            // computed against the stale max (P <= 256, exact in the fp32 sum and harmless in bf16) and O / l need no
            // correction.  The result is unchanged because O and l always share the same reference max.
            const float m_cand = fmaxf(m_run, mx * LOG2E);
            float alpha = 1.0f;
            if (m_cand > m_run + 8.0f) {                               // first valid block: m_run = -inf -> alpha = 0
                alpha = ex2(m_run - m_cand);
                m_run = m_cand;
            }
            const float m_use = (m_run == -CUDART_INF_F) ? 0.f : m_run;
            // exp2(s*log2e - m) and the row sum with packed fp32x2 FMA / ADD (FFMA2 / FADD2): half the issue slots
            const uint64_t sc2 = pack_f32x2(LOG2E, LOG2E), nm2 = pack_f32x2(-m_use, -m_use);
            uint64_t sum2[2] = {pack_f32x2(0.f, 0.f), pack_f32x2(0.f, 0.f)};
#pragma unroll
            for (int i = 0; i < ATT_BLOCK; i += 4) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    float x0, x1;
                    unpack_f32x2(fma_f32x2(pack_f32x2(s[i + 2 * u], s[i + 2 * u + 1]), sc2, nm2), x0, x1);
                    if (((i >> 1) + u) % 4 < POLY) {          // this pair goes to the FMA pipe
                        exp2_poly_pair(x0, x1);
                        s[i + 2 * u] = x0;
                        s[i + 2 * u + 1] = x1;
                    } else {                                  // this pair goes to the MUFU
                        s[i + 2 * u] = ex2(x0);
                        s[i + 2 * u + 1] = ex2(x1);
                    }
                    sum2[u] = add_f32x2(sum2[u], pack_f32x2(s[i + 2 * u], s[i + 2 * u + 1]));
                }
            }
            float sa, sb, sc, sd;
            unpack_f32x2(sum2[0], sa, sb);
            unpack_f32x2(sum2[1], sc, sd);
            l_run = l_run * alpha + ((sa + sb) + (sc + sd));
            // PV(j-1) must have drained the P buffer and finished O before either is touched again
            if (j > 0) {
                mbar_wait(&bar_kv_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
            }
            // P -> smem, bf16, K-major with the 128-B swizzle: 16-B chunk c of row r lands at chunk (c ^ (r & 7))
#pragma unroll
            for (int a = 0; a < 2; ++a) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int e = a * 64 + c * 8;
                    uint4 u;
                    u.x = pack_bf16x2(s[e + 0], s[e + 1]);
                    u.y = pack_bf16x2(s[e + 2], s[e + 3]);
                    u.z = pack_bf16x2(s[e + 4], s[e + 5]);
                    u.w = pack_bf16x2(s[e + 6], s[e + 7]);
                    *reinterpret_cast<uint4*>(p_row + a * (ATT_BLOCK * 128) + ((c ^ sw) << 4)) = u;
                }
            }
            // rescale the running O accumulator (complete through block j-1, see the wait above)
            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
                for (int c = 0; c < D / 16; ++c) {
                    uint32_t o[16];
                    tmem_ld16(tmem_o + lane_addr + c * 16, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st16(tmem_o + lane_addr + c * 16, o);
                }
                tmem_st_wait();
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p_full);
        }
        // epilogue: O / l -> bf16 -> HBM
        mbar_wait(bar_o_full, 0);
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        const bool row_ok = q0 + r < k_tokens;
        __nv_bfloat16* orow = out + (row_base + q0 + r) * h + head * D;
#pragma unroll
        for (int c = 0; c < D / 16; ++c) {
            uint32_t o[16];
            tmem_ld16(tmem_o + lane_addr + c * 16, o);
            tmem_ld_wait();
            if (row_ok) {
                uint4 u0, u1;
                u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l);
                u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l);
                u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l);
                u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l);
                u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l);
                u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l);
                u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l);
                u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l);
                reinterpret_cast<uint4*>(orow + c * 16)[0] = u0;
                reinterpret_cast<uint4*>(orow + c * 16)[1] = u1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ================================================================================================
// Persistent variant (default).  A CTA walks a static stride of work items (sequence, head, 128-query block) and
// treats all their KV blocks as ONE stream of iterations g = 0, 1, 2, ...:
//   * K/V ring, S / P hand-offs and barrier parities are indexed by g, so the loads for the next item's first blocks,
//     its Q tile (double-buffered) and its first S = QK^T are issued while the current item is still in softmax;
//   * P never touches shared memory: the softmax warps write it as packed bf16 into TMEM (tcgen05.st) and O += P V
//     takes its A operand from TMEM (tcgen05.mma ..., [a_tmem], b_desc).  At head_dim 64 the smem-P form spent
//     ~45 % of the shared-memory bandwidth of a KV block on writing P and reading it back, and shared memory
//     (UMMA operand reads + TMA fills), not the tensor pipe, is what bounds this kernel.
// TMEM columns: S [0,128) fp32 | P [128,192) bf16x2 | O [192, 192+D) fp32.
// ================================================================================================
struct AttnCursor {              // position in the flattened (item, kv-block) stream of this CTA
    int item;                    // global work-item index (n, head, qb); >= total -> stream exhausted
    int it;                      // local item counter of this CTA (0, 1, 2, ...)
    int j;                       // kv block inside the item
    int g;                       // global iteration counter
    int nkv, kvl, n, head, q0;
};

template <int D>
struct AttnPCfg {                // shared memory of the persistent kernel: Q x2 | K x3 | V x2 | barriers
    using B = AttnCfg<D>;
    // K and V have SEPARATE rings.  S(g+1) = Q K(g+1)^T is issued at the start of softmax(g), a whole block-time before
    // V(g+1) is needed, so a shared K/V stage (freed only by PV) made every K tile arrive one TMA latency late -- that
    // latency, not MUFU or shared memory, set the ~2000 cycles per KV block of the earlier versions.
    static constexpr int KST = 3, VST = 2;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = 2 * B::TILE_BYTES;
    static constexpr int OFF_V = (2 + KST) * B::TILE_BYTES;
    static constexpr int OFF_BAR = (2 + KST + VST) * B::TILE_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static constexpr int TMEM_COLS = (D == 128) ? 512 : 256;
    static constexpr int MIN_CTAS = SMEM_BYTES <= 115712 ? 2 : 1;       // (228 KB - 2 x 1 KB reserved) / 2
    static constexpr int TM_S = 0, TM_P = 128, TM_O = 192;
};

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, AttnPCfg<D>::MIN_CTAS)
attention_persistent_kernel(const __grid_constant__ CUtensorMap tma_qkv, int n_seq, int heads, int k_tokens, int h,
                            const int32_t* __restrict__ kv_info, const uint8_t* __restrict__ key_mask,
                            __nv_bfloat16* __restrict__ out, long long* __restrict__ dbg) {
    using Cfg = AttnCfg<D>;
    using PC = AttnPCfg<D>;
    // optional timeline of CTA 0 (tools/attn_timeline.py): dbg[role][iteration][slot] = clock64()
#define MOLLY_DBG(role, iter, slot)                                                             \
    do {                                                                                        \
        if (dbg != nullptr && blockIdx.x == 0 && (iter) < 64) dbg[((role) * 64 + (iter)) * 8 + (slot)] = clock64(); \
    } while (0)
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PC::OFF_BAR);
    uint64_t* bar_q = bars + 0;           // [2]  Q tile of item it landed in buffer it & 1
    uint64_t* bar_k_full = bars + 2;      // [3]
    uint64_t* bar_k_empty = bars + 5;     // [3]  S(g) complete: K stage g % 3 is free
    uint64_t* bar_v_full = bars + 8;      // [2]
    uint64_t* bar_v_empty = bars + 10;    // [2]  PV(g) complete: V stage g & 1 and P are free
    uint64_t* bar_s_full = bars + 12;
    uint64_t* bar_p_full = bars + 13;
    uint64_t* bar_s_free = bars + 14;
    uint64_t* bar_o_full = bars + 15;     // last PV of an item complete
    uint64_t* bar_o_free = bars + 16;     // softmax warps have read O out
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqb = (k_tokens + ATT_BLOCK - 1) / ATT_BLOCK;
    const int total = n_seq * heads * nqb;

    // cursor helpers (every role walks the same stream; items whose sequence is all padding are skipped)
    auto load_item = [&](AttnCursor& c) {
        while (c.item < total) {
            const int qb = c.item % nqb, t = c.item / nqb;
            c.head = t % heads;
            c.n = t / heads;
            c.q0 = qb * ATT_BLOCK;
            c.kvl = __ldg(kv_info + 2 * c.n);
            c.nkv = (c.kvl + ATT_BLOCK - 1) / ATT_BLOCK;
            if (c.nkv > 0) return;
            c.item += gridDim.x;              // all-pad sequence: no stream entries (softmax role zero-fills its rows)
        }
    };
    auto advance = [&](AttnCursor& c) {
        ++c.g;
        if (++c.j == c.nkv) {
            c.j = 0;
            ++c.it;
            c.item += gridDim.x;
            load_item(c);
        }
    };
    auto next_item = [&](AttnCursor& c) {     // item-granular step (Q prefetch cursor)
        ++c.it;
        c.item += gridDim.x;
        load_item(c);
    };
    auto first = [&]() {
        AttnCursor c;
        c.item = blockIdx.x; c.it = 0; c.j = 0; c.g = 0; c.nkv = 0; c.kvl = 0; c.n = 0; c.head = 0; c.q0 = 0;
        load_item(c);
        return c;
    };

    if (warp == 4) {
        if (lane == 0) {
            if ((smem_u32(smem) & 1023u) != 0) { printf("molly attention: smem base not 1024-B aligned\n"); __trap(); }
            tma_prefetch_desc(&tma_qkv);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bar_q[i], 1);
                mbar_init(&bar_v_full[i], 1);
                mbar_init(&bar_v_empty[i], 1);
            }
            for (int i = 0; i < PC::KST; ++i) {
                mbar_init(&bar_k_full[i], 1);
                mbar_init(&bar_k_empty[i], 1);
            }
            mbar_init(bar_s_full, 1);
            mbar_init(bar_p_full, 128);
            mbar_init(bar_s_free, 128);
            mbar_init(bar_o_full, 1);
            mbar_init(bar_o_free, 128);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, PC::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base + PC::TM_S;
    const uint32_t tmem_p = tmem_base + PC::TM_P;
    const uint32_t tmem_o = tmem_base + PC::TM_O;

    if (warp == 4) {
        if (lane == 0) {
            // ---------------- control thread: TMA producer + MMA issuer ----------------
            constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BLOCK, ATT_BLOCK, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BLOCK, D, false, true);      // B (= V) is MN-major
            const uint32_t s_q = smem_u32(smem + PC::OFF_Q), s_k = smem_u32(smem + PC::OFF_K);
            const uint32_t s_v = smem_u32(smem + PC::OFF_V);
            auto load_tile = [&](int smem_off, uint64_t* bar, int col, int row) {
#pragma unroll
                for (int b = 0; b < Cfg::NBOX; ++b)
                    tma_load_2d(smem + smem_off + b * Cfg::BOX_BYTES, &tma_qkv, bar, col + b * Cfg::BOX_D, row);
            };
            auto load_q = [&](const AttnCursor& c) {
                mbar_arrive_expect_tx(&bar_q[c.it & 1], Cfg::TILE_BYTES);
                load_tile(PC::OFF_Q + (c.it & 1) * Cfg::TILE_BYTES, &bar_q[c.it & 1], c.head * D, c.n * k_tokens + c.q0);
            };
            auto load_k = [&](const AttnCursor& c) {      // K(g) -> stage g % 3, once S(g-3) has released it
                const int stg = c.g % PC::KST;
                if (c.g >= PC::KST) mbar_wait(&bar_k_empty[stg], ((c.g / PC::KST) - 1) & 1);
                mbar_arrive_expect_tx(&bar_k_full[stg], Cfg::TILE_BYTES);
                load_tile(PC::OFF_K + stg * Cfg::TILE_BYTES, &bar_k_full[stg], h + c.head * D,
                          c.n * k_tokens + c.j * ATT_BLOCK);
            };
            auto load_v = [&](const AttnCursor& c) {      // V(g) -> stage g & 1, once PV(g-2) has released it
                const int stg = c.g & 1;
                if (c.g >= 2) mbar_wait(&bar_v_empty[stg], ((c.g >> 1) - 1) & 1);
                mbar_arrive_expect_tx(&bar_v_full[stg], Cfg::TILE_BYTES);
                load_tile(PC::OFF_V + stg * Cfg::TILE_BYTES, &bar_v_full[stg], 2 * h + c.head * D,
                          c.n * k_tokens + c.j * ATT_BLOCK);
            };
            // Base descriptors are built once; per MMA only the start-address field moves (one 32-bit add).
            const uint64_t qd0 = make_smem_desc(s_q, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
            const uint64_t kd0 = make_smem_desc(s_k, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
            const uint64_t vd0 = make_smem_desc(s_v, Cfg::BOX_BYTES, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
            auto issue_s = [&](const AttnCursor& c) {       // S = Q K^T : K-major x K-major, D/16 k-steps
                if (c.j == 0) mbar_wait(&bar_q[c.it & 1], (c.it >> 1) & 1);
                const int kst = c.g % PC::KST;
                mbar_wait(&bar_k_full[kst], (c.g / PC::KST) & 1);
                tc_fence_after();
                const uint64_t qd = desc_advance(qd0, (c.it & 1) * Cfg::TILE_BYTES);
                const uint64_t kd = desc_advance(kd0, kst * Cfg::TILE_BYTES);
#pragma unroll
                for (int s = 0; s < D / 16; ++s) {
                    const uint32_t off = ((s * 16) / Cfg::BOX_D) * Cfg::BOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    umma_bf16_ss(tmem_s, desc_advance(qd, off), desc_advance(kd, off), idesc_s, s != 0);
                }
                umma_commit(bar_s_full);
                umma_commit(&bar_k_empty[kst]);              // the same completion also releases the K stage
            };

            AttnCursor ck = first(), cv = ck, cs = ck, cp = ck, cq = ck;   // K-load / V-load / S-issue / PV / Q-load cursors
            if (cp.item < total) {
                load_q(cq); next_item(cq);
                for (int i = 0; i < PC::KST && ck.item < total; ++i) { load_k(ck); advance(ck); }
                for (int i = 0; i < PC::VST && cv.item < total; ++i) { load_v(cv); advance(cv); }
                if (cq.item < total) { load_q(cq); next_item(cq); }           // second Q buffer
                issue_s(cs); advance(cs);
                while (cp.item < total) {
                    const int g = cp.g, st = g & 1;
                    // (1) softmax holds S(g) in registers -> S(g+1) runs under softmax(g)
                    MOLLY_DBG(0, g, 0);
                    mbar_wait(bar_s_free, g & 1);
                    MOLLY_DBG(0, g, 1);
                    tc_fence_after();
                    if (cs.item < total) {
                        // S(g) was the last user of the previous item's Q buffer when cs starts a new item:
                        // refill that buffer with the Q tile of the item after cs
                        const bool new_item = cs.j == 0;
                        issue_s(cs);
                        if (new_item && cq.item < total && cq.it == cs.it + 1) { load_q(cq); next_item(cq); }
                        advance(cs);
                    }
                    // S(g) is complete (the softmax warps have read it), so its K stage is free: fetch K(g+3) now,
                    // two block-times before S(g+3) is issued
                    if (ck.item < total) { load_k(ck); advance(ck); }
                    MOLLY_DBG(0, g, 2);
                    // (2) O (+)= P(g) V(g) : P from TMEM (64 packed columns), V MN-major from smem; 8 k-steps of 16 keys
                    mbar_wait(bar_p_full, g & 1);
                    if (cp.j == 0 && cp.it >= 1) mbar_wait(bar_o_free, (cp.it - 1) & 1);   // previous item's O was read out
                    MOLLY_DBG(0, g, 3);
                    mbar_wait(&bar_v_full[st], (g >> 1) & 1);
                    tc_fence_after();
                    const uint64_t vd = desc_advance(vd0, st * Cfg::TILE_BYTES);
#pragma unroll
                    for (int s = 0; s < ATT_BLOCK / 16; ++s)
                        umma_bf16_ts(tmem_o, tmem_p + s * 8, desc_advance(vd, s * 16 * Cfg::ROW_BYTES), idesc_pv,
                                     (cp.j | s) != 0);
                    umma_commit(&bar_v_empty[st]);           // PV(g) done: V stage st and P are free
                    if (cp.j == cp.nkv - 1) umma_commit(bar_o_full);
                    MOLLY_DBG(0, g, 4);
                    // (3) refill V stage st with V(g+2) once PV(g) has drained it (needed two block-times from now)
                    if (cv.item < total) { load_v(cv); advance(cv); }
                    MOLLY_DBG(0, g, 5);
                    advance(cp);
                }
            }
        }
    } else {
        // ---------------- softmax warps: thread r owns query row r ----------------
        const int r = threadIdx.x;
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        int g = 0, it = 0;                                    // stream position, identical to the control thread's
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const int qb = item % nqb, tq = item / nqb;
            const int head = tq % heads, n = tq / heads;
            const int q0 = qb * ATT_BLOCK;
            const int kvl = __ldg(kv_info + 2 * n);
            const int nkv = (kvl + ATT_BLOCK - 1) / ATT_BLOCK;
            const long long row_base = static_cast<long long>(n) * k_tokens;
            if (nkv == 0) {          // all-pad sequence: not in the stream; zero-fill (the reference never encodes one)
                if (q0 + r < k_tokens) {
                    uint4* o = reinterpret_cast<uint4*>(out + (row_base + q0 + r) * h + head * D);
                    for (int i = 0; i < D / 8; ++i) o[i] = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            const bool interior = __ldg(kv_info + 2 * n + 1) != kvl;       // pad ids before the last real token
            float m_run = -CUDART_INF_F;      // running reference max, log2 domain
            float l_run = 0.f;
            for (int j = 0; j < nkv; ++j, ++g) {
                if (threadIdx.x == 0) MOLLY_DBG(1, g, 0);
                mbar_wait(bar_s_full, g & 1);
                if (threadIdx.x == 0) MOLLY_DBG(1, g, 1);
                tc_fence_after();
                float s[ATT_BLOCK];
                {
                    uint32_t raw[ATT_BLOCK];
                    tmem_ld32(tmem_s + lane_addr + 0, raw);
                    tmem_ld32(tmem_s + lane_addr + 32, raw + 32);
                    tmem_ld32(tmem_s + lane_addr + 64, raw + 64);
                    tmem_ld32(tmem_s + lane_addr + 96, raw + 96);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < ATT_BLOCK; ++i) s[i] = __uint_as_float(raw[i]);
                }
                tc_fence_before();
                mbar_arrive(bar_s_free);                     // S(g) is in registers: the MMA warp may start S(g+1)
                if (threadIdx.x == 0) MOLLY_DBG(1, g, 2);
                const int j0 = j * ATT_BLOCK;
                if (interior) {
                    const uint4* mk = reinterpret_cast<const uint4*>(key_mask + row_base + j0);
                    const bool vec_ok = ((row_base + j0) & 15) == 0 && j0 + ATT_BLOCK <= k_tokens;
#pragma unroll
                    for (int gq = 0; gq < 8; ++gq) {
                        uint32_t w[4];
                        if (vec_ok) {
                            const uint4 u = __ldg(mk + gq);
                            w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                w[i] = 0;
#pragma unroll
                                for (int b = 0; b < 4; ++b) {
                                    const int cc = j0 + gq * 16 + i * 4 + b;
                                    const uint32_t v = (cc < k_tokens) ? key_mask[row_base + cc] : 0;
                                    w[i] |= (v & 0xffu) << (8 * b);
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const bool ok = ((w[i >> 2] >> (8 * (i & 3))) & 0xffu) != 0 && (j0 + gq * 16 + i < kvl);
                            if (!ok) s[gq * 16 + i] = -CUDART_INF_F;
                        }
                    }
                } else if (j0 + ATT_BLOCK > kvl) {
                    const int lim = kvl - j0;
#pragma unroll
                    for (int i = 0; i < ATT_BLOCK; ++i)
                        if (i >= lim) s[i] = -CUDART_INF_F;
                }
                float mx4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
                for (int i = 4; i < ATT_BLOCK; i += 4) {
                    mx4[0] = fmaxf(mx4[0], s[i]); mx4[1] = fmaxf(mx4[1], s[i + 1]);
                    mx4[2] = fmaxf(mx4[2], s[i + 2]); mx4[3] = fmaxf(mx4[3], s[i + 3]);
                }
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                // lazy rescaling (see the non-persistent kernel): move the reference max only when it grew by > 2^8
                const float m_cand = fmaxf(m_run, mx * LOG2E);
                float alpha = 1.0f;
                if (m_cand > m_run + 8.0f) {
                    alpha = ex2(m_run - m_cand);
                    m_run = m_cand;
                }
                const float m_use = (m_run == -CUDART_INF_F) ? 0.f : m_run;
                const uint64_t sc2 = pack_f32x2(LOG2E, LOG2E), nm2 = pack_f32x2(-m_use, -m_use);
                uint64_t sum2[2] = {pack_f32x2(0.f, 0.f), pack_f32x2(0.f, 0.f)};
                // The warp issues in order: an add placed right behind its two MUFU producers stalls for the MUFU latency
                // and idles the XU pipe.  The row sum therefore trails the exponentials by SUM_LAG pairs.
                constexpr int SUM_LAG = 6;
#pragma unroll
                for (int i = 0; i < ATT_BLOCK / 2 + SUM_LAG; ++i) {
                    if (i < ATT_BLOCK / 2) {
                        float x0, x1;
                        unpack_f32x2(fma_f32x2(pack_f32x2(s[2 * i], s[2 * i + 1]), sc2, nm2), x0, x1);
                        s[2 * i] = ex2(x0);
                        s[2 * i + 1] = ex2(x1);
                    }
                    if (i >= SUM_LAG) {
                        const int k = i - SUM_LAG;
                        sum2[k & 1] = add_f32x2(sum2[k & 1], pack_f32x2(s[2 * k], s[2 * k + 1]));
                    }
                }
                float sa, sb, sc, sd;
                unpack_f32x2(sum2[0], sa, sb);
                unpack_f32x2(sum2[1], sc, sd);
                l_run = l_run * alpha + ((sa + sb) + (sc + sd));
                if (threadIdx.x == 0) MOLLY_DBG(1, g, 3);
                // PV(g-1) must have consumed P (and, inside an item, finished O) before either is touched again
                if (g > 0) {
                    mbar_wait(&bar_v_empty[(g - 1) & 1], ((g - 1) >> 1) & 1);
                    tc_fence_after();
                }
                // P -> TMEM as packed bf16 pairs: column c of lane r holds keys (2c, 2c+1) of query row r
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {             // 16 packed columns at a time keeps the register peak low
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2_alu(s[hh * 32 + 2 * i], s[hh * 32 + 2 * i + 1]);
                    tmem_st16(tmem_p + lane_addr + hh * 16, pk);
                }
                if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
                    for (int cc = 0; cc < D / 16; ++cc) {
                        uint32_t o[16];
                        tmem_ld16(tmem_o + lane_addr + cc * 16, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st16(tmem_o + lane_addr + cc * 16, o);
                    }
                }
                if (threadIdx.x == 0) MOLLY_DBG(1, g, 4);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_p_full);
                if (threadIdx.x == 0) MOLLY_DBG(1, g, 5);
            }
            // item epilogue: O / l -> bf16 -> HBM, then hand O back to the MMA thread
            mbar_wait(bar_o_full, it & 1);
            tc_fence_after();
            const float inv_l = 1.0f / l_run;
            const bool row_ok = q0 + r < k_tokens;
            __nv_bfloat16* orow = out + (row_base + q0 + r) * h + head * D;
#pragma unroll
            for (int cc = 0; cc < D / 16; ++cc) {
                uint32_t o[16];
                tmem_ld16(tmem_o + lane_addr + cc * 16, o);
                tmem_ld_wait();
                if (row_ok) {
                    uint4 u0, u1;
                    u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l);
                    u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l);
                    u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l);
                    u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l);
                    u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l);
                    u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l);
                    u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l);
                    u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l);
                    reinterpret_cast<uint4*>(orow + cc * 16)[0] = u0;
                    reinterpret_cast<uint4*>(orow + cc * 16)[1] = u1;
                }
            }
            tc_fence_before();
            mbar_arrive(bar_o_free);
            ++it;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, PC::TMEM_COLS);
    }
}

long long* g_attn_debug = nullptr;     // set through molly_attention_debug(); nullptr in production

bool attention_persistent_enabled() {
    static int v = -1;
    if (v < 0) {
        // Measured (tools/attn_bench.py, ESM-650M layer shape): one item per CTA with P staged in shared memory 579 us,
        // persistent + smem P 586 us, persistent + P in TMEM 666 us -> the simple kernel is the default;
        // MOLLY_ATTN_PERSISTENT=1 selects the streamed, P-in-TMEM variant.
        const char* e = getenv("MOLLY_ATTN_PERSISTENT");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

template <int D>
int launch_attention_persistent(const CUtensorMap& tm, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                                const uint8_t* key_mask, void* out, cudaStream_t stream) {
    using Cfg = AttnPCfg<D>;
    auto kernel = attention_persistent_kernel<D>;
    static bool configured = false;
    if (!configured) {
        MOLLY_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int total = n_seq * heads * ((k_tokens + ATT_BLOCK - 1) / ATT_BLOCK);
    const int slots = device_sm_count() * Cfg::MIN_CTAS;
    const int grid = total < slots ? total : slots;
    {   // dense-equivalent work 4*n*K*K*h (exact when every sequence is full length)
        ProfScope prof(PF_ATTENTION, 4.0 * n_seq * k_tokens * static_cast<double>(k_tokens) * h, stream);
        kernel<<<grid, ATT_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, n_seq, heads, k_tokens, h, kv_info, key_mask,
                                                               static_cast<__nv_bfloat16*>(out), g_attn_debug);
    }
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

template <int D>
int launch_attention(const CUtensorMap& tm, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                     const uint8_t* key_mask, void* out, cudaStream_t stream) {
    if (attention_persistent_enabled())
        return launch_attention_persistent<D>(tm, n_seq, k_tokens, h, heads, kv_info, key_mask, out, stream);
    using Cfg = AttnCfg<D>;
    static int poly = -1;             // pairs out of 4 whose exp2 runs on the FMA pipe (MOLLY_ATTN_POLY = 0 | 1 | 2)
    if (poly < 0) {
        const char* e = getenv("MOLLY_ATTN_POLY");
        poly = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : ATTN_POLY_DEFAULT;
    }
    auto kernel = poly == 0 ? attention_kernel<D, 0> : (poly == 1 ? attention_kernel<D, 1> : attention_kernel<D, 2>);
    static bool configured = false;
    if (!configured) {
        MOLLY_CUDA(cudaFuncSetAttribute(attention_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention_kernel<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    dim3 grid((k_tokens + ATT_BLOCK - 1) / ATT_BLOCK, heads, n_seq);
    {   // dense-equivalent work 4*n*K*K*h (exact when every sequence is full length)
        ProfScope prof(PF_ATTENTION, 4.0 * n_seq * k_tokens * static_cast<double>(k_tokens) * h, stream);
        kernel<<<grid, ATT_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, k_tokens, h, kv_info, key_mask,
                                                               static_cast<__nv_bfloat16*>(out));
    }
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

}  // namespace

void attention_set_debug(long long* buf) { g_attn_debug = buf; }

int attention_make_map(CUtensorMap* tq, const void* qkv, int rows, int h, int heads) {
    const int d = h / heads;
    MOLLY_CHECK(d == 16 || d == 32 || d == 64 || d == 128, MOLLY_ERR_UNSUPPORTED,
                "attention: head_dim %d not in {16,32,64,128}", d);
    const int box_d = d < 64 ? d : 64;
    return make_tma_2d(tq, qkv, rows, 3 * h, 3 * h, ATT_BLOCK, box_d, 2);
}

int attention_launch(const CUtensorMap& tqkv, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                     const uint8_t* key_mask, void* out, cudaStream_t stream) {
    MOLLY_CHECK(n_seq > 0 && k_tokens > 0 && heads > 0 && h % heads == 0, MOLLY_ERR_INVALID,
                "attention: bad shape n_seq=%d k=%d h=%d heads=%d", n_seq, k_tokens, h, heads);
    MOLLY_CHECK(static_cast<long long>(n_seq) * k_tokens < (1ll << 31), MOLLY_ERR_UNSUPPORTED,
                "attention: n_seq*k_tokens exceeds int32 TMA coordinates");
    MOLLY_CHECK(n_seq <= 65535 && heads <= 65535, MOLLY_ERR_UNSUPPORTED, "attention: grid too large");
    switch (h / heads) {
        case 16: return launch_attention<16>(tqkv, n_seq, k_tokens, h, heads, kv_info, key_mask, out, stream);
        case 32: return launch_attention<32>(tqkv, n_seq, k_tokens, h, heads, kv_info, key_mask, out, stream);
        case 64: return launch_attention<64>(tqkv, n_seq, k_tokens, h, heads, kv_info, key_mask, out, stream);
        case 128: return launch_attention<128>(tqkv, n_seq, k_tokens, h, heads, kv_info, key_mask, out, stream);
        default: MOLLY_CHECK(false, MOLLY_ERR_UNSUPPORTED, "attention: head_dim %d unsupported", h / heads);
    }
}

}  // namespace molly
