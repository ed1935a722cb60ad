// Fused bidirectional multi-head attention, second generation (head_dim <= 64): ONE CTA per SM that runs TWO independent
// 128-query-row pipelines ("tiles"), each the stream-of-work-items pipeline of attention.cu, restructured around what the
// round-2 timelines showed (tools/attn2_timeline.py, profiles/r02_attention.md): a softmax warp that owns whole 128-key rows
// runs its exp2 stream at 12.5 cycles per MUFU.EX2 alone (8 is the pipe's rate) and spends as long again outside it (wait
// for S, tcgen05.ld, row max, hand-offs, item epilogue), and with two such warps per scheduler running in lock-step the MUFU
// idles ~45 % of the time.  A single warp's instruction stream, not a pipe, was the bound -- so this kernel puts FOUR softmax
// warps on every scheduler:
//   * 20 warps: softmax warps 0-7 (tile 0) and 8-15 (tile 1); TWO threads per query row -- warps w and w+4 of a tile share
//     TMEM lane quarter w % 4, the first takes keys 0..63 of every KV block, the second keys 64..127.  The row max goes
//     through shared memory (one float + a 64-thread named barrier per block); the row sums stay per thread and are added
//     in the item epilogue; each thread rescales / writes out half of the O columns.
//     Warps 16 / 17 lane 0 = the tiles' control threads (TMA producer + tcgen05.mma issuer); warp 18 allocates TMEM.
//     `setmaxnreg`: control group 96 -> 64, softmax warps 96 -> 104 registers.
//   * P never touches shared memory: packed bf16 pairs go straight into TMEM (tcgen05.st) and O += P V takes its A operand
//     from tensor memory -- no STS.128 + fence.proxy.async on the softmax chain, no P traffic on the shared-memory pipe.
//   * row max with the 3-input FMNMX3; P stored in 32-key chunks as the exponentials retire.
// TMEM columns of tile t (base 256 t): S [0,128) fp32 | P [128,192) bf16x2 | O [192, 192 + D) fp32.
// Semantics are those of attention.cu (HF:257-282 eager attention with the key-padding mask of HF:679-709).
#include <math_constants.h>
#include <stdlib.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

void attention2_set_debug(long long*) {}

namespace {

constexpr int A2_BLOCK = 128;                  // query rows per tile == keys per KV block
constexpr int A2_THREADS = 640;                // 16 softmax warps + the control warp-group
constexpr int A2_REGS_SOFTMAX = 104;
constexpr int A2_REGS_CONTROL = 64;
constexpr int A2_HALF = A2_BLOCK / 2;          // keys per thread per KV block
constexpr float A2_LOG2E = 1.4426950408889634f;

template <int D, int NST>
struct A2Cfg {
    static_assert(D == 16 || D == 32 || D == 64, "attention2: head_dim 16 / 32 / 64");
    static constexpr int ROW_BYTES = D * 2;                                   // 32 / 64 / 128: one swizzle span
    static constexpr uint32_t LAYOUT = ROW_BYTES == 128 ? kLayoutSW128 : (ROW_BYTES == 64 ? kLayoutSW64 : kLayoutSW32);
    static constexpr int TILE_BYTES = A2_BLOCK * ROW_BYTES;                   // a Q, K or V tile (one TMA box)
    static constexpr int OFF_Q = 0;                                           // inside a tile's shared-memory slice
    static constexpr int OFF_K = TILE_BYTES;
    static constexpr int OFF_V = OFF_K + NST * TILE_BYTES;
    static constexpr int SLICE_BYTES = (1 + 2 * NST) * TILE_BYTES;
    static constexpr int OFF_BAR = 2 * SLICE_BYTES;
    static constexpr int NBAR = 5 + 2 * NST;                                  // per tile
    static constexpr int OFF_X = OFF_BAR + 2 * NBAR * 8 + 16;                 // row-max / row-sum exchange of the half-row threads
    static constexpr int X_FLOATS = 2 * 2 * A2_BLOCK + 2 * A2_BLOCK;          // per tile: max [g parity][half][row], sum [half][row]
    static constexpr int SMEM_BYTES = OFF_X + 2 * X_FLOATS * 4;
    static constexpr int TM_S = 0, TM_P = 128, TM_O = 192, TM_TILE = 256;
};

struct A2Bars {                                // one tile's barriers
    uint64_t* q;
    uint64_t* kv_full;                         // [NST]
    uint64_t* kv_empty;                        // [NST]  PV(g) complete: stage g % NST, P and O are free
    uint64_t* s_full;
    uint64_t* p_full;                          // 256 arrivals
    uint64_t* o_full;
    uint64_t* s_free;                          // 256 arrivals: S(g) is in registers
};
template <int NST>
__device__ __forceinline__ A2Bars a2_bars(uint64_t* base) {
    A2Bars b;
    b.q = base;
    b.kv_full = base + 1;
    b.kv_empty = base + 1 + NST;
    b.s_full = base + 1 + 2 * NST;
    b.p_full = base + 2 + 2 * NST;
    b.o_full = base + 3 + 2 * NST;
    b.s_free = base + 4 + 2 * NST;
    return b;
}

struct A2Item { int n, head, q0, kvl, nkv, n_nonpad; };
struct A2Shape { int n_seq, heads, k_tokens, h, nqb, total; };

__device__ __forceinline__ A2Item a2_decode(int item, const A2Shape& sh, const int32_t* __restrict__ kv_info) {
    A2Item w;                                  // query block fastest: neighbouring tiles share one (sequence, head)'s K/V in L2
    w.q0 = (item % sh.nqb) * A2_BLOCK;
    w.head = (item / sh.nqb) % sh.heads;
    w.n = item / (sh.nqb * sh.heads);
    w.kvl = kv_info[2 * w.n];
    w.n_nonpad = kv_info[2 * w.n + 1];
    w.nkv = (w.kvl + A2_BLOCK - 1) / A2_BLOCK;
    return w;
}

__device__ __forceinline__ float a2_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float a2_max3(float a, float b, float c) {          // FMNMX3
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// exp2 on the FMA pipe: 2^x = 2^n p(r), n = round(x), r = x - n in [-0.5, 0.5], cubic p (max relative error 1.0e-4, below the
// bf16 rounding of P that follows), 2^n spliced into the exponent field.  Same polynomial as attention.cu.
__device__ __forceinline__ void a2_exp2_poly_pair(float& x0, float& x1) {
    const uint64_t magic = pack_f32x2(12582912.0f, 12582912.0f);
    const uint64_t x2 = pack_f32x2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
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

// ------------------------------------------------------------------------------------------------------------------
// control thread of one tile: TMA producer + MMA issuer over the tile's stream of work items
// ------------------------------------------------------------------------------------------------------------------
template <int D, int NST>
__device__ __forceinline__ void a2_control(uint8_t* slice, const A2Bars& bar, uint32_t tmem_tile, const CUtensorMap* tma_qkv,
                                           int slot, int stride, const A2Shape& sh, const int32_t* __restrict__ kv_info) {
    using Cfg = A2Cfg<D, NST>;
    constexpr uint32_t idesc_s = make_idesc_bf16(A2_BLOCK, A2_BLOCK, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(A2_BLOCK, D, false, true);             // B (= V) is MN-major
    const uint32_t s_q = smem_u32(slice + Cfg::OFF_Q), s_k = smem_u32(slice + Cfg::OFF_K), s_v = smem_u32(slice + Cfg::OFF_V);
    const uint32_t tmem_s = tmem_tile + Cfg::TM_S, tmem_p = tmem_tile + Cfg::TM_P, tmem_o = tmem_tile + Cfg::TM_O;
    const int total = sh.total, h = sh.h, k_tokens = sh.k_tokens;

    auto next_item = [&](int item) {                         // first item >= `item` of this tile that has keys
        while (item < total && kv_info[2 * (item / (sh.nqb * sh.heads))] <= 0) item += stride;
        return item;
    };
    auto load_q = [&](const A2Item& w) {
        mbar_arrive_expect_tx(bar.q, Cfg::TILE_BYTES);
        tma_load_2d(slice + Cfg::OFF_Q, tma_qkv, bar.q, w.head * D, w.n * k_tokens + w.q0);
    };
    // load cursor: runs NST KV blocks ahead of the compute cursor, across item boundaries
    int l_item = next_item(slot), l_j = 0, g_load = 0;
    A2Item lw = a2_decode(l_item < total ? l_item : 0, sh, kv_info);
    auto load_next_kv = [&]() {                              // stream block g_load -> stage g_load % NST (caller: stage is free)
        if (l_item >= total) return;
        const int stg = g_load % NST, row = lw.n * k_tokens + l_j * A2_BLOCK;
        mbar_arrive_expect_tx(&bar.kv_full[stg], 2 * Cfg::TILE_BYTES);
        tma_load_2d(slice + Cfg::OFF_K + stg * Cfg::TILE_BYTES, tma_qkv, &bar.kv_full[stg], h + lw.head * D, row);
        tma_load_2d(slice + Cfg::OFF_V + stg * Cfg::TILE_BYTES, tma_qkv, &bar.kv_full[stg], 2 * h + lw.head * D, row);
        ++g_load;
        if (++l_j == lw.nkv) {
            l_item = next_item(l_item + stride);
            l_j = 0;
            if (l_item < total) lw = a2_decode(l_item, sh, kv_info);
        }
    };
    auto issue_s = [&](int g) {                              // S = Q K(g)^T : K-major x K-major, D/16 k-steps
        const int st = g % NST;
        mbar_wait(&bar.kv_full[st], (g / NST) & 1);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < D / 16; ++s) {
            const uint64_t qd = make_smem_desc(s_q + s * 32, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
            const uint64_t kd = make_smem_desc(s_k + st * Cfg::TILE_BYTES + s * 32, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
            umma_bf16_ss(tmem_s, qd, kd, idesc_s, s != 0);
        }
        umma_commit(bar.s_full);
    };

    int c_item = l_item, it = 0, g = 0;
    if (c_item >= total) return;
    load_q(lw);
#pragma unroll
    for (int i = 0; i < NST; ++i) load_next_kv();
    mbar_wait(bar.q, 0);
    issue_s(0);
    while (c_item < total) {
        const A2Item w = a2_decode(c_item, sh, kv_info);
        const int nxt = next_item(c_item + stride);
        for (int j = 0; j < w.nkv; ++j, ++g) {
            const int st = g % NST;
            const bool last = j == w.nkv - 1;
            // (1) the moment the softmax warps hold S(g) in registers, S(g+1) is issued: it runs under softmax(g).
            //     On an item's last block Q is dead instead: the next item's Q is fetched into the same buffer.
            mbar_wait(bar.s_free, g & 1);
            tc_fence_after();
            if (!last) issue_s(g + 1);
            else if (nxt < total) load_q(a2_decode(nxt, sh, kv_info));
            // (2) O += P(g) V(g) : P from tensor memory (lane = query row, 8 packed columns per 16 keys), V MN-major from smem
            mbar_wait(bar.p_full, g & 1);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < A2_BLOCK / 16; ++s) {
                const uint64_t vd = make_smem_desc(s_v + st * Cfg::TILE_BYTES + s * 16 * Cfg::ROW_BYTES, Cfg::TILE_BYTES,
                                                   8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
                umma_bf16_ts(tmem_o, tmem_p + s * 8, vd, idesc_pv, (j | s) != 0);
            }
            umma_commit(&bar.kv_empty[st]);                  // PV(g) done: stage st, P and O are free
            if (last) {
                umma_commit(bar.o_full);
                if (nxt < total) {                           // first S of the next item, under this item's epilogue
                    mbar_wait(bar.q, (it + 1) & 1);
                    issue_s(g + 1);
                }
            }
            // (3) refill stage st with stream block g + NST once PV(g) has drained it
            if (l_item < total) {
                mbar_wait(&bar.kv_empty[st], (g / NST) & 1);
                load_next_kv();
            }
        }
        ++it;
        c_item = nxt;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// softmax warps of one tile: two threads per query row r; `half` selects this thread's 64 keys of every KV block
// ------------------------------------------------------------------------------------------------------------------
template <int D, int NST, int POLY>
__device__ __forceinline__ void a2_softmax(const A2Bars& bar, uint32_t tmem_tile, float* xch, int r, int half, int pair_bar,
                                           int slot, int stride, const A2Shape& sh, const int32_t* __restrict__ kv_info,
                                           const uint8_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out,
                                           float* __restrict__ lse2) {
    using Cfg = A2Cfg<D, NST>;
    constexpr int OC = D / 2;                                                 // O columns this thread rescales / writes out
    const uint32_t lane_addr = static_cast<uint32_t>(r & ~31) << 16;          // TMEM lane quarter of this warp
    const uint32_t tmem_s = tmem_tile + Cfg::TM_S + lane_addr + half * A2_HALF;
    const uint32_t tmem_p = tmem_tile + Cfg::TM_P + lane_addr + half * (A2_HALF / 2);
    const uint32_t tmem_o = tmem_tile + Cfg::TM_O + lane_addr + half * OC;
    float* const xmax = xch;                                                  // [2 (g parity)][2 (half)][128]
    float* const xsum = xch + 2 * 2 * A2_BLOCK;                               // [2 (half)][128]
    const int k_tokens = sh.k_tokens, h = sh.h;
    int it = 0, g = 0;
    for (int item = slot; item < sh.total; item += stride) {
        const A2Item w = a2_decode(item, sh, kv_info);
        const int kvl = w.kvl;
        const bool interior = w.n_nonpad != w.kvl;                            // pad ids before the last real token
        const long long row_base = static_cast<long long>(w.n) * k_tokens;
        const bool row_ok = w.q0 + r < k_tokens;
        __nv_bfloat16* orow = out + (row_base + w.q0 + r) * h + w.head * D + half * OC;
        if (w.nkv == 0) {                                                     // all-pad sequence: the reference never encodes one
            if (row_ok) {
                for (int i = 0; i < OC / 8; ++i) reinterpret_cast<uint4*>(orow)[i] = make_uint4(0, 0, 0, 0);
                if (lse2 != nullptr && half == 0)
                    lse2[(static_cast<size_t>(w.n) * sh.heads + w.head) * k_tokens + w.q0 + r] = -CUDART_INF_F;
            }
            continue;
        }
        float m_run = -CUDART_INF_F;                                          // running reference max, log2 domain (both halves agree)
        float l_run = 0.f;                                                    // row sum over THIS thread's keys
        for (int j = 0; j < w.nkv; ++j, ++g) {
            const int j0 = j * A2_BLOCK + half * A2_HALF;                     // first key of this thread in the block
            mbar_wait(bar.s_full, g & 1);
            tc_fence_after();
            float s[A2_HALF];
            {
                uint32_t raw[A2_HALF];
                tmem_ld32(tmem_s, raw);
                tmem_ld32(tmem_s + 32, raw + 32);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < A2_HALF; ++i) s[i] = __uint_as_float(raw[i]);
            }
            tc_fence_before();
            mbar_arrive(bar.s_free);                                          // the control thread may issue S(g+1)
            if (interior) {
                const uint4* mk = reinterpret_cast<const uint4*>(key_mask + row_base + j0);
                const bool vec_ok = ((row_base + j0) & 15) == 0 && j0 + A2_HALF <= k_tokens;
#pragma unroll
                for (int q = 0; q < A2_HALF / 16; ++q) {
                    uint32_t wd[4];
                    if (vec_ok) {
                        const uint4 u = __ldg(mk + q);
                        wd[0] = u.x; wd[1] = u.y; wd[2] = u.z; wd[3] = u.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            wd[i] = 0;
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const int c = j0 + q * 16 + i * 4 + b;
                                const uint32_t v = (c < k_tokens) ? key_mask[row_base + c] : 0;
                                wd[i] |= (v & 0xffu) << (8 * b);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const bool ok = ((wd[i >> 2] >> (8 * (i & 3))) & 0xffu) != 0 && (j0 + q * 16 + i < kvl);
                        if (!ok) s[q * 16 + i] = -CUDART_INF_F;
                    }
                }
            } else if (j0 + A2_HALF > kvl) {
                const int lim = kvl - j0;
#pragma unroll
                for (int i = 0; i < A2_HALF; ++i)
                    if (i >= lim) s[i] = -CUDART_INF_F;
            }
            // row max of this thread's 64 keys: four independent FMNMX3 chains, then the other half's through shared memory
            float mx4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
            for (int i = 4; i + 8 <= A2_HALF; i += 8) {
                mx4[0] = a2_max3(mx4[0], s[i], s[i + 1]);
                mx4[1] = a2_max3(mx4[1], s[i + 2], s[i + 3]);
                mx4[2] = a2_max3(mx4[2], s[i + 4], s[i + 5]);
                mx4[3] = a2_max3(mx4[3], s[i + 6], s[i + 7]);
            }
            mx4[0] = a2_max3(mx4[0], s[A2_HALF - 4], s[A2_HALF - 3]);
            mx4[1] = a2_max3(mx4[1], s[A2_HALF - 2], s[A2_HALF - 1]);
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            {
                float* slot_x = xmax + (g & 1) * (2 * A2_BLOCK);
                slot_x[half * A2_BLOCK + r] = mx;
                named_bar_sync(pair_bar, 64);                                 // the two warps that share this lane quarter
                mx = fmaxf(mx, slot_x[(half ^ 1) * A2_BLOCK + r]);
            }
            // Lazy rescaling (see attention.cu): the reference max moves only when the row max grew by more than 2^8, so
            // P <= 256 (exact in the fp32 sum, harmless in bf16) and O / l almost never need a correction.
            const float m_cand = fmaxf(m_run, mx * A2_LOG2E);
            float alpha = 1.0f;
            if (m_cand > m_run + 8.0f) {                                      // first valid block: m_run = -inf -> alpha = 0
                alpha = a2_ex2(m_run - m_cand);
                m_run = m_cand;
            }
            const float m_use = (m_run == -CUDART_INF_F) ? 0.f : m_run;
            const uint64_t sc2 = pack_f32x2(A2_LOG2E, A2_LOG2E), nm2 = pack_f32x2(-m_use, -m_use);
            uint64_t sum2[2] = {pack_f32x2(0.f, 0.f), pack_f32x2(0.f, 0.f)};
            // P = exp2(s log2e - m) in 32-key chunks: packed bf16 pairs go to TMEM as each chunk retires
            // (column c of lane r holds keys 2c, 2c+1 of query row r: the K-major A operand of O += P V)
#pragma unroll
            for (int c = 0; c < A2_HALF / 32; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x0, x1;
                    unpack_f32x2(fma_f32x2(pack_f32x2(s[c * 32 + 2 * i], s[c * 32 + 2 * i + 1]), sc2, nm2), x0, x1);
                    if (i % 4 < POLY) {                                       // this pair goes to the FMA pipe
                        a2_exp2_poly_pair(x0, x1);
                    } else {                                                  // this pair goes to the MUFU
                        x0 = a2_ex2(x0);
                        x1 = a2_ex2(x1);
                    }
                    sum2[i & 1] = add_f32x2(sum2[i & 1], pack_f32x2(x0, x1));
                    pk[i] = pack_bf16x2(x0, x1);
                }
                if (c == 0 && j > 0) {                                        // PV(g-1) consumed P and finished O
                    mbar_wait(&bar.kv_empty[(g - 1) % NST], ((g - 1) / NST) & 1);     // (j == 0: the previous item's o_full)
                    tc_fence_after();
                }
                tmem_st16(tmem_p + c * 16, pk);
            }
            float sa, sb, sc, sd;
            unpack_f32x2(sum2[0], sa, sb);
            unpack_f32x2(sum2[1], sc, sd);
            l_run = l_run * alpha + ((sa + sb) + (sc + sd));
            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {            // rare: rescale this thread's half of the O columns
#pragma unroll
                for (int c = 0; c < OC / 8; ++c) {
                    uint32_t o[8];
                    tmem_ld8(tmem_o + c * 8, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st8(tmem_o + c * 8, o);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar.p_full);
        }
        // item epilogue: the row sum is the two halves' sums; O / l -> bf16 -> HBM, each thread half of the columns
        xsum[half * A2_BLOCK + r] = l_run;
        named_bar_sync(pair_bar, 64);
        const float l_tot = l_run + xsum[(half ^ 1) * A2_BLOCK + r];
        mbar_wait(bar.o_full, it & 1);
        ++it;
        tc_fence_after();
        const float inv_l = 1.0f / l_tot;
        if (lse2 != nullptr && row_ok && half == 0)                           // row log-sum-exp, log2 domain (backward)
            lse2[(static_cast<size_t>(w.n) * sh.heads + w.head) * k_tokens + w.q0 + r] = m_run + log2f(l_tot);
#pragma unroll
        for (int c = 0; c < OC / 8; ++c) {
            uint32_t o[8];
            tmem_ld8(tmem_o + c * 8, o);
            tmem_ld_wait();
            if (row_ok) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l);
                u.y = pack_bf16x2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l);
                u.z = pack_bf16x2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l);
                u.w = pack_bf16x2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l);
                reinterpret_cast<uint4*>(orow)[c] = u;
            }
        }
        tc_fence_before();                                                    // O is read: the next item's PV may overwrite it
        named_bar_sync(pair_bar, 64);                                         // xsum is read: the next item may overwrite it
    }
}

template <int D, int NST, int POLY>
__global__ void __launch_bounds__(A2_THREADS, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tma_qkv, int n_seq, int heads, int k_tokens, int h,
                  const int32_t* __restrict__ kv_info, const uint8_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out,
                  float* __restrict__ lse2) {
    using Cfg = A2Cfg<D, NST>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::NBAR);
    float* xch = reinterpret_cast<float*>(smem + Cfg::OFF_X);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    A2Shape sh;
    sh.n_seq = n_seq; sh.heads = heads; sh.k_tokens = k_tokens; sh.h = h;
    sh.nqb = (k_tokens + A2_BLOCK - 1) / A2_BLOCK;
    sh.total = n_seq * heads * sh.nqb;

    if (warp == 18) {
        if (lane == 0) {
            if ((smem_u32(smem) & 1023u) != 0) { printf("molly attention2: smem base not 1024-B aligned\n"); __trap(); }
            tma_prefetch_desc(&tma_qkv);
            for (int t = 0; t < 2; ++t) {
                const A2Bars b = a2_bars<NST>(bars + t * Cfg::NBAR);
                mbar_init(b.q, 1);
                for (int st = 0; st < NST; ++st) { mbar_init(&b.kv_full[st], 1); mbar_init(&b.kv_empty[st], 1); }
                mbar_init(b.s_full, 1);
                mbar_init(b.p_full, 256);
                mbar_init(b.o_full, 1);
                mbar_init(b.s_free, 256);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int stride = 2 * gridDim.x;

    if (warp >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A2_REGS_CONTROL));
        if (warp < 18 && lane == 0) {
            const int t = warp - 16;
            a2_control<D, NST>(smem + t * Cfg::SLICE_BYTES, a2_bars<NST>(bars + t * Cfg::NBAR), tmem_base + t * Cfg::TM_TILE,
                               &tma_qkv, 2 * blockIdx.x + t, stride, sh, kv_info);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(A2_REGS_SOFTMAX));
        const int t = warp >> 3, half = (warp >> 2) & 1;
        a2_softmax<D, NST, POLY>(a2_bars<NST>(bars + t * Cfg::NBAR), tmem_base + t * Cfg::TM_TILE, xch + t * Cfg::X_FLOATS,
                                 threadIdx.x & 127, half, 1 + t * 4 + (warp & 3), 2 * blockIdx.x + t, stride, sh, kv_info,
                                 key_mask, out, lse2);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 18) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int a2_env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e == nullptr ? dflt : atoi(e);
}

template <int D, int NST>
int launch_attention2(const CUtensorMap& tm, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                      const uint8_t* key_mask, void* out, float* lse, cudaStream_t stream) {
    using Cfg = A2Cfg<D, NST>;
    static int poly = -1;             // pairs out of 4 whose exp2 runs on the FMA pipe (MOLLY_ATTN_POLY = 0 | 1 | 2)
    if (poly < 0) {
        poly = a2_env_int("MOLLY_ATTN_POLY", ATTENTION2_POLY_DEFAULT);
        if (poly < 0 || poly > 2) poly = ATTENTION2_POLY_DEFAULT;
    }
    auto kernel = poly == 0 ? attention2_kernel<D, NST, 0> : (poly == 1 ? attention2_kernel<D, NST, 1> : attention2_kernel<D, NST, 2>);
    static bool configured = false;
    if (!configured) {
        MOLLY_CUDA(cudaFuncSetAttribute(attention2_kernel<D, NST, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention2_kernel<D, NST, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention2_kernel<D, NST, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int total = n_seq * heads * ((k_tokens + A2_BLOCK - 1) / A2_BLOCK);
    const int sms = device_sm_count();
    const int grid = (total + 1) / 2 < sms ? (total + 1) / 2 : sms;
    {
        prof_attention_work(kv_info, n_seq, k_tokens, h, 4.0, stream);      // work = 4 h K sum(kv_len), known on the device only
        ProfScope prof(PF_ATTENTION, 4.0 * n_seq * k_tokens * static_cast<double>(k_tokens) * h, stream);
        kernel<<<grid, A2_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, n_seq, heads, k_tokens, h, kv_info, key_mask,
                                                              static_cast<__nv_bfloat16*>(out), lse);
    }
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

}  // namespace

bool attention2_enabled(int d) {      // MOLLY_ATTN_V2 = 0 | 1 overrides the default
    static int v = -1;
    if (v < 0) v = a2_env_int("MOLLY_ATTN_V2", ATTENTION2_DEFAULT) != 0 ? 1 : 0;
    return v == 1 && d <= 64;
}

int attention2_launch(const AttnMaps& maps, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                      const uint8_t* key_mask, void* out, float* lse, cudaStream_t stream) {
    switch (h / heads) {
        case 16: return launch_attention2<16, 2>(maps.q, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        case 32: return launch_attention2<32, 2>(maps.q, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        case 64: return launch_attention2<64, 2>(maps.q, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        default: MOLLY_CHECK(false, MOLLY_ERR_UNSUPPORTED, "attention2: head_dim %d unsupported", h / heads);
    }
}

}  // namespace molly
