// Fused bidirectional multi-head attention, second generation (head_dim <= 64): ONE CTA per SM that runs TWO independent
// 128-query-row pipelines ("tiles") over streams of work items (sequence, head, 128-query block), FlashAttention-4 style:
//   * 16 warps.  Softmax warps 0-3 (tile 0) and 4-7 (tile 1): thread r owns query row r of its tile and does nothing but
//     S -> row max -> exp2 -> P.  Epilogue warps 8-11: O / l -> bf16 -> HBM of BOTH tiles (and the zero rows of all-pad
//     sequences), so the softmax warps never wait for the item's last O += P V, never read O, never store (the round-2
//     timeline of the first version of this kernel: ~4 000 cycles per work item went there).  Warps 12 / 13 lane 0: the
//     tiles' control threads (TMA producer + tcgen05.mma issuer); warp 14 allocates the 512 TMEM columns.
//     `setmaxnreg`: softmax 128 -> 200 registers, epilogue and control groups 128 -> 56.
//   * the exp2 phases of the two softmax warps that share a scheduler are SEQUENCED (a token per warp pair, passed through
//     an mbarrier): the timelines showed both tiles entering their exp2 phase together -- each warp then gets half of the
//     MUFU -- and leaving it together, after which the MUFU idles while both do their serial work (wait for S, tcgen05.ld,
//     row max, hand-offs).  Taking turns, one warp's exp2 phase runs at the full MUFU rate under the other's serial work.
//   * P never touches shared memory: packed bf16 pairs go straight into TMEM (tcgen05.st) and O += P V takes its A operand
//     from tensor memory; row max with the 3-input FMNMX3; P stored in 32-key chunks as the exponentials retire.
// TMEM columns of tile t (base 256 t): S [0,128) fp32 | P [128,192) bf16x2 | O [192, 192 + D) fp32.
// Semantics are those of attention.cu (HF:257-282 eager attention with the key-padding mask of HF:679-709).
#include <math_constants.h>
#include <stdlib.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

// -DA2_TIMELINE: clock64 stamps of the softmax thread r == 0 of both tiles of the first 8 CTAs, 8 slots per stream block
// (tools/attn2_timeline.py): 0 block start, 1 S seen, 2 S in registers, 3 row max done, 4 exp2 turn acquired,
// 5 exp2 done / turn passed on, 6 P stored + p_full arrived.
#ifdef A2_TIMELINE
__device__ long long* d_a2_tl = nullptr;
void attention2_set_debug(long long* buf) { cudaMemcpyToSymbol(d_a2_tl, &buf, sizeof(buf)); }
#define TL2(slot) do { if (tl_on) tl_buf[tl_base + (slot)] = clock64(); } while (0)
#else
void attention2_set_debug(long long*) {}
#define TL2(slot) do { } while (0)
#endif

namespace {

constexpr int A2_BLOCK = 128;                  // query rows per tile == keys per KV block
constexpr int A2_THREADS = 512;                // 8 softmax warps + the epilogue warp-group + the control warp-group
constexpr int A2_REGS_SOFTMAX = 200;
constexpr int A2_REGS_OTHER = 56;
#ifndef A2_SEQUENCE
#define A2_SEQUENCE 1                          // 0: the tiles' exp2 phases are not sequenced (A/B build)
#endif
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
    static constexpr int NBAR = 7 + 2 * NST;                                  // per tile
    static constexpr int OFF_SEQ = OFF_BAR + 2 * NBAR * 8;                    // [2 tiles][4 lane quarters] exp2-turn barriers
    static constexpr int OFF_DONE = OFF_SEQ + 8 * 8;                          // [2][4] int: that softmax warp has left its stream
    static constexpr int OFF_SLOT = OFF_DONE + 8 * 4;                         // TMEM base
    static constexpr int OFF_STATS = (OFF_SLOT + 16 + 15) / 16 * 16;          // [2 tiles][2 item parities][128 rows] (l, m)
    static constexpr int SMEM_BYTES = OFF_STATS + 2 * 2 * A2_BLOCK * 8;
    static constexpr int TM_S = 0, TM_P = 128, TM_O = 192, TM_TILE = 256;
};

struct A2Bars {                                // one tile's barriers
    uint64_t* q;
    uint64_t* kv_full;                         // [NST]
    uint64_t* kv_empty;                        // [NST]  PV(g) complete: stage g % NST, P and O are free
    uint64_t* s_full;
    uint64_t* p_full;                          // 128 arrivals
    uint64_t* o_full;                          // the item's last O += P V completed
    uint64_t* s_free;                          // 128 arrivals: S(g) is in registers
    uint64_t* stats_full;                      // 128 arrivals: the item's row sums / maxima are in shared memory
    uint64_t* o_free;                          // 128 arrivals: the epilogue warps have read the item's O
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
    b.stats_full = base + 5 + 2 * NST;
    b.o_free = base + 6 + 2 * NST;
    return b;
}

struct A2Item { int n, head, q0, kvl, nkv, n_nonpad; };
struct A2Shape { int n_seq, heads, k_tokens, h, nqb, total; };

// work-item slot of tile t of this CTA and the stride of the item streams (-DA2_ONE_TILE: tile 1 idle, a bring-up experiment)
__device__ __forceinline__ int a2_slot(int t, int total) {
#ifdef A2_ONE_TILE
    return t == 0 ? static_cast<int>(blockIdx.x) : total;
#else
    (void)total;
    return 2 * static_cast<int>(blockIdx.x) + t;
#endif
}
__device__ __forceinline__ int a2_stride() {
#ifdef A2_ONE_TILE
    return static_cast<int>(gridDim.x);
#else
    return 2 * static_cast<int>(gridDim.x);
#endif
}

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
            if (j == 0 && it > 0) mbar_wait(bar.o_free, (it - 1) & 1);       // the epilogue warps have read the previous item's O
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
                if (nxt < total) {                           // first S of the next item
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
// softmax warps of one tile: thread r owns query row r
// ------------------------------------------------------------------------------------------------------------------
// exp2-turn token of the warp pair (tile 0 warp q, tile 1 warp q) that shares scheduler q: `mine` completes a phase when the
// other tile's warp has finished an exp2 phase (32 arrivals); tile 0 starts with the token (tile 1's warps pre-arrive).
struct A2Turn {
    uint64_t* mine;
    uint64_t* theirs;
    volatile int* peer_done;
    volatile int* my_done;
    int n;                                      // exp2 phases this warp has waited for
    bool solo;                                  // the other tile's warp has left its stream: no more turns
};
__device__ __forceinline__ void a2_turn_begin(A2Turn& t) {
#if A2_SEQUENCE
    if (t.solo) return;
    const uint32_t parity = t.n & 1;
    ++t.n;
    if (mbar_try_wait(t.mine, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(t.mine, parity)) {
        // the "other warp has left" flag is probed rarely: a shared-memory load per spin is an LDS stream next to the other
        // tile's exp2 loop on the same scheduler, which slows that loop 2.5x (tools/ubench/interfere.cu); try_wait costs nothing
        if ((++spins & 63u) == 0) {
            if (*t.peer_done) { t.solo = true; return; }
            if (clock64() - t0 > MOLLY_MBAR_TIMEOUT_CYCLES) {
                printf("molly attention2: exp2-turn timeout block=%d thread=%d\n", blockIdx.x, threadIdx.x);
                __trap();
            }
        }
    }
#endif
}
__device__ __forceinline__ void a2_turn_end(A2Turn& t) {
#if A2_SEQUENCE
    if (!t.solo) mbar_arrive(t.theirs);
#endif
}

template <int D, int NST, int POLY>
__device__ __forceinline__ void a2_softmax(const A2Bars& bar, uint32_t tmem_tile, float2* stats, A2Turn turn, int r, int slot,
                                           int stride, const A2Shape& sh, const int32_t* __restrict__ kv_info,
                                           const uint8_t* __restrict__ key_mask) {
    using Cfg = A2Cfg<D, NST>;
    const uint32_t lane_addr = static_cast<uint32_t>(r & ~31) << 16;          // TMEM lane quarter of this warp
    const uint32_t tmem_s = tmem_tile + Cfg::TM_S + lane_addr, tmem_p = tmem_tile + Cfg::TM_P + lane_addr;
    const uint32_t tmem_o = tmem_tile + Cfg::TM_O + lane_addr;
    const int k_tokens = sh.k_tokens;
#ifdef A2_TIMELINE
    long long* const tl_buf = d_a2_tl;                                        // read once: a stamp must not cost a global load
#endif
    int it = 0, g = 0;
    for (int item = slot; item < sh.total; item += stride) {
        const A2Item w = a2_decode(item, sh, kv_info);
        if (w.nkv == 0) continue;                                             // all-pad sequence: the epilogue warps write its zeros
        const int kvl = w.kvl;
        const bool interior = w.n_nonpad != w.kvl;                            // pad ids before the last real token
        const long long row_base = static_cast<long long>(w.n) * k_tokens;
        float m_run = -CUDART_INF_F;                                          // running reference max, log2 domain
        float l_run = 0.f;
        for (int j = 0; j < w.nkv; ++j, ++g) {
            const int j0 = j * A2_BLOCK;
#ifdef A2_TIMELINE
            const bool tl_on = tl_buf != nullptr && r == 0 && blockIdx.x < 8 && g < 64;
            const int tl_base = ((blockIdx.x * 2 + ((tmem_tile >> 8) & 1)) * 64 + g) * 8;
#endif
            TL2(0);
            mbar_wait(bar.s_full, g & 1);
            TL2(1);
            tc_fence_after();
            float s[A2_BLOCK];
            {
                uint32_t raw[A2_BLOCK];
#pragma unroll
                for (int c = 0; c < A2_BLOCK / 32; ++c) tmem_ld32(tmem_s + c * 32, raw + c * 32);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < A2_BLOCK; ++i) s[i] = __uint_as_float(raw[i]);
            }
            tc_fence_before();
            mbar_arrive(bar.s_free);                                          // the control thread may issue S(g+1)
            TL2(2);
            if (interior) {
                const uint4* mk = reinterpret_cast<const uint4*>(key_mask + row_base + j0);
                const bool vec_ok = ((row_base + j0) & 15) == 0 && j0 + A2_BLOCK <= k_tokens;
#pragma unroll
                for (int q = 0; q < A2_BLOCK / 16; ++q) {
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
            } else if (j0 + A2_BLOCK > kvl) {
                const int lim = kvl - j0;
#pragma unroll
                for (int i = 0; i < A2_BLOCK; ++i)
                    if (i >= lim) s[i] = -CUDART_INF_F;
            }
            // row max: four independent FMNMX3 chains
            float mx4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
            for (int i = 4; i + 8 <= A2_BLOCK; i += 8) {
                mx4[0] = a2_max3(mx4[0], s[i], s[i + 1]);
                mx4[1] = a2_max3(mx4[1], s[i + 2], s[i + 3]);
                mx4[2] = a2_max3(mx4[2], s[i + 4], s[i + 5]);
                mx4[3] = a2_max3(mx4[3], s[i + 6], s[i + 7]);
            }
            mx4[0] = a2_max3(mx4[0], s[A2_BLOCK - 4], s[A2_BLOCK - 3]);
            mx4[1] = a2_max3(mx4[1], s[A2_BLOCK - 2], s[A2_BLOCK - 1]);
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // Lazy rescaling (see attention.cu): the reference max moves only when the row max grew by more than 2^8, so
            // P <= 256 (exact in the fp32 sum, harmless in bf16) and O / l almost never need a correction.
            const float m_cand = fmaxf(m_run, mx * A2_LOG2E);
            float alpha = 1.0f;
            if (m_cand > m_run + 8.0f) {                                      // first valid block: m_run = -inf -> alpha = 0
                alpha = a2_ex2(m_run - m_cand);
                m_run = m_cand;
            }
            const float m_use = (m_run == -CUDART_INF_F) ? 0.f : m_run;
            // The exp2 phase proper holds the scheduler's MUFU exclusively (turn-taking), so everything that is not a
            // MUFU.EX2 or the pack of its result stays OUT of it: the scale-and-shift FFMA2s run before the turn is taken
            // (volatile: they must not sink below the wait), the wait for PV(g-1) too, and the row sum (FADD2) runs after
            // the turn has been passed on, from the exponentials kept in place in s[].
            {
                const uint64_t sc2 = pack_f32x2(A2_LOG2E, A2_LOG2E), nm2 = pack_f32x2(-m_use, -m_use);
#pragma unroll
                for (int i = 0; i < A2_BLOCK; i += 2) {
                    uint64_t v;
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(pack_f32x2(s[i], s[i + 1])), "l"(sc2), "l"(nm2));
                    unpack_f32x2(v, s[i], s[i + 1]);
                }
            }
            if (g > 0) {                                                      // PV(g-1) consumed P and finished O
                mbar_wait(&bar.kv_empty[(g - 1) % NST], ((g - 1) / NST) & 1);
                tc_fence_after();
            }
            TL2(3);
            a2_turn_begin(turn);                                              // this warp's turn on the scheduler's MUFU
            TL2(4);
            // P = exp2(.), packed bf16 pairs go to TMEM in 32-key chunks (column c of lane r holds keys 2c, 2c+1 of query row r:
            // the K-major A operand of O += P V).  The MUFU stream runs A2_AHEAD key pairs ahead of the pack stream, in an
            // order pinned with volatile asm: left to itself ptxas packs a pair right behind its two MUFU.EX2, and the
            // in-order warp then waits out the MUFU latency on every pair (measured: 14 cycles per MUFU.EX2 instead of 8).
            {
                constexpr int A2_AHEAD = 8, NP = A2_BLOCK / 2;
                auto exp_pair = [&](int p) {
                    float x0 = s[2 * p], x1 = s[2 * p + 1];
                    if (p % 4 < POLY) {                                       // this pair goes to the FMA pipe
                        a2_exp2_poly_pair(x0, x1);
                    } else {                                                  // this pair goes to the MUFU
                        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x0) : "f"(x0));
                        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x1) : "f"(x1));
                    }
                    s[2 * p] = x0;
                    s[2 * p + 1] = x1;
                };
#pragma unroll
                for (int p = 0; p < A2_AHEAD; ++p) exp_pair(p);
                uint32_t pk[16];
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    if (p + A2_AHEAD < NP) exp_pair(p + A2_AHEAD);
                    if (p + A2_AHEAD == NP - 1) {
                        a2_turn_end(turn);                                    // last MUFU.EX2 issued: the other tile's warp may start
                        TL2(5);
                    }
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[p % 16]) : "f"(s[2 * p + 1]), "f"(s[2 * p]));
                    if (p % 16 == 15) tmem_st16(tmem_p + (p / 16) * 16, pk);
                }
            }
            uint64_t sum2[2] = {pack_f32x2(0.f, 0.f), pack_f32x2(0.f, 0.f)};
#pragma unroll
            for (int i = 0; i < A2_BLOCK; i += 4) {
                sum2[0] = add_f32x2(sum2[0], pack_f32x2(s[i], s[i + 1]));
                sum2[1] = add_f32x2(sum2[1], pack_f32x2(s[i + 2], s[i + 3]));
            }
            float sa, sb, sc, sd;
            unpack_f32x2(sum2[0], sa, sb);
            unpack_f32x2(sum2[1], sc, sd);
            l_run = l_run * alpha + ((sa + sb) + (sc + sd));
            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {            // rare: rescale the running O accumulator
#pragma unroll
                for (int c = 0; c < D / 16; ++c) {
                    uint32_t o[16];
                    tmem_ld16(tmem_o + c * 16, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st16(tmem_o + c * 16, o);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar.p_full);
            TL2(6);
        }
        // the item's row statistics go to the epilogue warps (buffer it & 1: the epilogue of item it-2 finished before the
        // control thread issued the first O += P V of item it-1, see o_free)
        stats[(it & 1) * A2_BLOCK + r] = make_float2(l_run, m_run);
        mbar_arrive(bar.stats_full);                                          // (release: the store above is visible to the waiter)
        ++it;
    }
#if A2_SEQUENCE
    // leave the turn-taking: the other tile's warp must not wait for a token that will never come
    __syncwarp();
    if ((r & 31) == 0) *turn.my_done = 1;
    __threadfence_block();
    a2_turn_end(turn);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// epilogue warps: O / l -> bf16 -> HBM for the work items of BOTH tiles, in the order the tiles finish them
// ------------------------------------------------------------------------------------------------------------------
template <int D, int NST>
__device__ __forceinline__ void a2_epilogue(uint64_t* bars, uint32_t tmem_base, const float2* stats_all, int r,
                                            const A2Shape& sh, const int32_t* __restrict__ kv_info,
                                            __nv_bfloat16* __restrict__ out, float* __restrict__ lse2) {
    using Cfg = A2Cfg<D, NST>;
    const uint32_t lane_addr = static_cast<uint32_t>(r & ~31) << 16;
    const int k_tokens = sh.k_tokens, h = sh.h;
    const int stride = a2_stride();
    // The two tiles finish their items at their own pace (different kv_len), and with the exp2 turn-taking a tile that is
    // held up holds up the other one: the items are therefore served in the order they become READY (non-blocking probes),
    // never in a fixed tile order -- waiting for tile 0 while tile 1's finished item keeps its control thread from issuing
    // the next O += P V would deadlock the pair.
    int item[2] = {a2_slot(0, sh.total), a2_slot(1, sh.total)};
    int it[2] = {0, 0};
    A2Item w[2];
    auto seek = [&](int t) {                                                  // next item of tile t that has keys; zero rows for the rest
        while (item[t] < sh.total) {
            w[t] = a2_decode(item[t], sh, kv_info);
            if (w[t].nkv > 0) return;
            if (w[t].q0 + r < k_tokens) {                                     // all-pad sequence: the reference never encodes one
                __nv_bfloat16* orow = out + (static_cast<long long>(w[t].n) * k_tokens + w[t].q0 + r) * h + w[t].head * D;
                for (int i = 0; i < D / 8; ++i) reinterpret_cast<uint4*>(orow)[i] = make_uint4(0, 0, 0, 0);
                if (lse2 != nullptr)
                    lse2[(static_cast<size_t>(w[t].n) * sh.heads + w[t].head) * k_tokens + w[t].q0 + r] = -CUDART_INF_F;
            }
            item[t] += stride;
        }
    };
    seek(0);
    seek(1);
    long long t_idle = clock64();
    while (item[0] < sh.total || item[1] < sh.total) {
        bool served = false;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (item[t] >= sh.total) continue;
            const A2Bars bar = a2_bars<NST>(bars + t * Cfg::NBAR);
            // warp-uniform readiness (every lane must take the same branch: tcgen05.ld is warp-wide)
            const bool ready = mbar_test_wait(bar.stats_full, it[t] & 1) && mbar_test_wait(bar.o_full, it[t] & 1);
            if (!__all_sync(0xffffffffu, ready)) continue;
            served = true;
            const bool row_ok = w[t].q0 + r < k_tokens;
            __nv_bfloat16* orow = out + (static_cast<long long>(w[t].n) * k_tokens + w[t].q0 + r) * h + w[t].head * D;
            float* lrow = lse2 == nullptr ? nullptr
                                          : lse2 + (static_cast<size_t>(w[t].n) * sh.heads + w[t].head) * k_tokens + w[t].q0 + r;
            const uint32_t tmem_o = tmem_base + t * Cfg::TM_TILE + Cfg::TM_O + lane_addr;
            const float2 st = stats_all[(t * 2 + (it[t] & 1)) * A2_BLOCK + r];
            ++it[t];
            tc_fence_after();
            const float inv_l = 1.0f / st.x;
            if (row_ok && lrow != nullptr) *lrow = st.y + log2f(st.x);        // row log-sum-exp, log2 domain (backward)
#pragma unroll
            for (int c = 0; c < D / 16; ++c) {                                // 16 columns at a time: this group has 56 registers
                uint32_t o[16];
                tmem_ld16(tmem_o + c * 16, o);
                tmem_ld_wait();
                uint4 u0, u1;
                u0.x = pack_bf16x2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l);
                u0.y = pack_bf16x2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l);
                u0.z = pack_bf16x2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l);
                u0.w = pack_bf16x2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l);
                u1.x = pack_bf16x2(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l);
                u1.y = pack_bf16x2(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l);
                u1.z = pack_bf16x2(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l);
                u1.w = pack_bf16x2(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l);
                if (c == D / 16 - 1) {                                        // O is in registers: the next item's first PV may go
                    tc_fence_before();
                    mbar_arrive(bar.o_free);
                }
                if (row_ok) {
                    reinterpret_cast<uint4*>(orow + c * 16)[0] = u0;
                    reinterpret_cast<uint4*>(orow + c * 16)[1] = u1;
                }
            }
            item[t] += stride;
            seek(t);
        }
        if (served) {
            t_idle = clock64();
        } else {
            __nanosleep(200);
            if (clock64() - t_idle > MOLLY_MBAR_TIMEOUT_CYCLES) {
                printf("molly attention2: epilogue timeout block=%d thread=%d\n", blockIdx.x, threadIdx.x);
                __trap();
            }
        }
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
    uint64_t* seq = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_SEQ);
    int* done = reinterpret_cast<int*>(smem + Cfg::OFF_DONE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_SLOT);
    float2* stats = reinterpret_cast<float2*>(smem + Cfg::OFF_STATS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    A2Shape sh;
    sh.n_seq = n_seq; sh.heads = heads; sh.k_tokens = k_tokens; sh.h = h;
    sh.nqb = (k_tokens + A2_BLOCK - 1) / A2_BLOCK;
    sh.total = n_seq * heads * sh.nqb;

    if (warp == 14) {
        if (lane == 0) {
            if ((smem_u32(smem) & 1023u) != 0) { printf("molly attention2: smem base not 1024-B aligned\n"); __trap(); }
            tma_prefetch_desc(&tma_qkv);
            for (int t = 0; t < 2; ++t) {
                const A2Bars b = a2_bars<NST>(bars + t * Cfg::NBAR);
                mbar_init(b.q, 1);
                for (int st = 0; st < NST; ++st) { mbar_init(&b.kv_full[st], 1); mbar_init(&b.kv_empty[st], 1); }
                mbar_init(b.s_full, 1);
                mbar_init(b.p_full, 128);
                mbar_init(b.o_full, 1);
                mbar_init(b.s_free, 128);
                mbar_init(b.stats_full, 128);
                mbar_init(b.o_free, 128);
            }
            for (int i = 0; i < 8; ++i) { mbar_init(&seq[i], 32); done[i] = 0; }
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
    const int stride = a2_stride();

    if (warp >= 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A2_REGS_OTHER));
        if (warp < 14 && elect_one()) {      // (elect.sync, not a lane test: see the control thread of attention.cu)
            const int t = warp - 12;
            a2_control<D, NST>(smem + t * Cfg::SLICE_BYTES, a2_bars<NST>(bars + t * Cfg::NBAR), tmem_base + t * Cfg::TM_TILE,
                               &tma_qkv, a2_slot(t, sh.total), stride, sh, kv_info);
        }
    } else if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A2_REGS_OTHER));
        a2_epilogue<D, NST>(bars, tmem_base, stats, threadIdx.x & 127, sh, kv_info, out, lse2);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(A2_REGS_SOFTMAX));
        const int t = warp >> 2, q = warp & 3;
        A2Turn turn;
        turn.mine = &seq[t * 4 + q];
        turn.theirs = &seq[(1 - t) * 4 + q];
        turn.my_done = &done[t * 4 + q];
        turn.peer_done = &done[(1 - t) * 4 + q];
        turn.n = 0;
        turn.solo = false;
#if A2_SEQUENCE
        if (t == 1) mbar_arrive(turn.theirs);                                 // tile 0 has the first turn
#endif
        a2_softmax<D, NST, POLY>(a2_bars<NST>(bars + t * Cfg::NBAR), tmem_base + t * Cfg::TM_TILE, stats + t * 2 * A2_BLOCK,
                                 turn, threadIdx.x & 127, a2_slot(t, sh.total), stride, sh, kv_info, key_mask);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 14) {
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
