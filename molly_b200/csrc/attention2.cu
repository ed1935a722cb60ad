// Fused bidirectional multi-head attention, second generation (head_dim <= 64): ONE CTA per SM that runs TWO independent
// 128-query-row pipelines ("tiles") over streams of work items (sequence, head, 128-query block), plus what one CTA per SM
// allows:
//   * 12 warps: softmax warps 0-3 (tile 0) and 4-7 (tile 1), thread r owns query row r of its tile; warp 8 / 9 lane 0 = the
//     tile's control thread (TMA producer + tcgen05.mma issuer); warp 10 allocates the 512 TMEM columns.  Every scheduler holds
//     exactly two softmax warps and one control-group warp, so `setmaxnreg` can move registers: control group 168 -> 72,
//     softmax warps 168 -> 216 (no spills; attention.cu's 2 CTAs x 5 warps were capped at 168).
//   * P never touches shared memory: the softmax warps write it as packed bf16 pairs straight into TMEM (tcgen05.st) and
//     O += P V takes its A operand from tensor memory -- no 16 x STS.128 + fence.proxy.async on the softmax warp's chain,
//     and no P write / P read on the shared-memory pipe.
//   * SOFTWARE-PIPELINED softmax.  ncu on the first version of this kernel (profiles/r02a_ncu.md): ptxas already paces the
//     exp2 loop at the MUFU rate (MUFU, MUFU, FADD2, F2FP every 16 cycles), yet the MUFU was only 55 % busy because half of
//     each softmax warp's time went to everything else (wait for S, tcgen05.ld, row max, hand-offs, item epilogue) while
//     its MUFU slots stayed empty.  Now the next block's S is pulled from TMEM in 32-key chunks into the registers the
//     current block's chunks leave behind, and its mask + row max (FMNMX3) are interleaved into the exp2 stream of the
//     current block, whose issue slots are 3/4 empty.  What stays serial per 128-key block is the exp2 loop itself plus
//     ~32 FMNMX3 and the two hand-offs.
//   * K ring (2 stages, freed when S = Q K^T has been read) and V ring (2 stages, freed when O += P V completed) are separate,
//     and Q is double-buffered per tile: S(g+1) is always issued the moment S(g) is in registers, across item boundaries.
// TMEM columns of tile t (base 256 t): S [0,128) fp32 | P [128,192) bf16x2 | O [192, 192 + D) fp32.
// Semantics are those of attention.cu (HF:257-282 eager attention with the key-padding mask of HF:679-709).
#include <math_constants.h>
#include <stdlib.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

// -DA2_TIMELINE: clock64 stamps of the softmax thread r == 0 of both tiles of the first 8 CTAs, 8 slots per stream block
// (tools/attn2_timeline.py): 0 step start, 1 keys 0..63 exponentiated, 2 PV(g-1) seen, 3 S(g+1) seen, 4 keys 64..95 done,
// 5 keys 96..127 done, 6 p_full arrived, 7 step end (next block's max, item epilogue).
#ifdef A2_TIMELINE
__device__ long long* d_a2_tl = nullptr;
void attention2_set_debug(long long* buf) { cudaMemcpyToSymbol(d_a2_tl, &buf, sizeof(buf)); }
#define TL2(slot) do { if (tl_on) d_a2_tl[tl_base + (slot)] = clock64(); } while (0)
#else
void attention2_set_debug(long long*) {}
#define TL2(slot) do { } while (0)
#endif

namespace {

constexpr int A2_BLOCK = 128;                  // query rows per tile == keys per KV block
constexpr int A2_THREADS = 384;                // 8 softmax warps + the control warp-group
constexpr int A2_REGS_SOFTMAX = 216;
constexpr int A2_REGS_CONTROL = 72;
constexpr float A2_LOG2E = 1.4426950408889634f;

template <int D>
struct A2Cfg {
    static_assert(D == 16 || D == 32 || D == 64, "attention2: head_dim 16 / 32 / 64");
    static constexpr int ROW_BYTES = D * 2;                                   // 32 / 64 / 128: one swizzle span
    static constexpr uint32_t LAYOUT = ROW_BYTES == 128 ? kLayoutSW128 : (ROW_BYTES == 64 ? kLayoutSW64 : kLayoutSW32);
    static constexpr int TILE_BYTES = A2_BLOCK * ROW_BYTES;                   // a Q, K or V tile (one TMA box)
    static constexpr int OFF_Q = 0;                                           // inside a tile's shared-memory slice: Q x2
    static constexpr int OFF_K = 2 * TILE_BYTES;                              // K x2
    static constexpr int OFF_V = 4 * TILE_BYTES;                              // V x2
    static constexpr int SLICE_BYTES = 6 * TILE_BYTES;
    static constexpr int OFF_BAR = 2 * SLICE_BYTES;
    static constexpr int NBAR = 12;                                           // per tile
    static constexpr int SMEM_BYTES = OFF_BAR + 2 * NBAR * 8 + 16;
    static constexpr int TM_S = 0, TM_P = 128, TM_O = 192, TM_TILE = 256;
};

struct A2Bars {                                // one tile's barriers
    uint64_t* q;                               // [2]  Q tile of item `it` landed in buffer it & 1
    uint64_t* k_full;                          // [2]
    uint64_t* v_full;                          // [2]
    uint64_t* pv_done;                         // [2]  O += P(g) V(g) complete: V stage g & 1, P and O are free
    uint64_t* s_full;
    uint64_t* p_full;                          // 128 arrivals
    uint64_t* o_full;
    uint64_t* s_free;                          // 128 arrivals: S(g) is in registers
};
__device__ __forceinline__ A2Bars a2_bars(uint64_t* base) {
    A2Bars b;
    b.q = base;
    b.k_full = base + 2;
    b.v_full = base + 4;
    b.pv_done = base + 6;
    b.s_full = base + 8;
    b.p_full = base + 9;
    b.o_full = base + 10;
    b.s_free = base + 11;
    return b;
}

struct A2Shape { int n_seq, heads, k_tokens, h, nqb, total; };

// Position in one tile's stream of KV blocks.  Both the control thread and the softmax threads walk it.
struct A2Cur {
    int item;                                  // global work item (>= total: stream exhausted)
    int it;                                    // ordinal of the item among this tile's items that have keys
    int j;                                     // KV block inside the item
    int g;                                     // ordinal of the block in the tile's stream
    int n, head, q0, kvl, nkv, n_nonpad;
};
__device__ __forceinline__ bool a2_valid(const A2Cur& c, const A2Shape& sh) { return c.item < sh.total; }
// decode c.item, skipping items without keys (all-pad sequences); on_empty(c) is called for each skipped item
template <typename F>
__device__ __forceinline__ void a2_seek(A2Cur& c, const A2Shape& sh, const int32_t* __restrict__ kv_info, int stride,
                                        F&& on_empty) {
    while (c.item < sh.total) {                // query block fastest: neighbouring tiles share one (sequence, head)'s K/V in L2
        c.q0 = (c.item % sh.nqb) * A2_BLOCK;
        c.head = (c.item / sh.nqb) % sh.heads;
        c.n = c.item / (sh.nqb * sh.heads);
        c.kvl = kv_info[2 * c.n];
        c.n_nonpad = kv_info[2 * c.n + 1];
        c.nkv = (c.kvl + A2_BLOCK - 1) / A2_BLOCK;
        if (c.nkv > 0) return;
        on_empty(c);
        c.item += stride;
    }
}
// the control thread's decode, out of line (five cursors advance through it; register-only interface):
// returns (item, n, head, q0) of the first work item >= `item` of this tile's stride that has keys (item >= total: none)
__device__ __forceinline__ int4 a2_seek_ctrl(int item, int total, int nqb, int heads, int stride,
                                          const int32_t* __restrict__ kv_info) {
    int4 r = make_int4(item, 0, 0, 0);
    while (r.x < total) {
        r.w = (r.x % nqb) * A2_BLOCK;
        r.z = (r.x / nqb) % heads;
        r.y = r.x / (nqb * heads);
        if (kv_info[2 * r.y] > 0) break;
        r.x += stride;
    }
    return r;
}
// The softmax threads keep two of these live next to ~170 data registers: only what every block needs; (n, head, q0) are
// re-derived from `item` once per item (epilogue) or in the rare interior-pad path.
struct A2Lite {
    int item;                                  // global work item (>= total: stream exhausted)
    int j, nkv, kvl;                           // KV block inside the item, #blocks, valid keys
    bool interior;                             // pad ids before the last real token: per-key byte mask needed
};
template <typename F>
__device__ __forceinline__ void a2_seek_lite(A2Lite& c, const A2Shape& sh, const int32_t* __restrict__ kv_info, int stride,
                                             F&& on_empty) {
    while (c.item < sh.total) {
        const int n = c.item / (sh.nqb * sh.heads);
        c.kvl = kv_info[2 * n];
        c.nkv = (c.kvl + A2_BLOCK - 1) / A2_BLOCK;
        if (c.nkv > 0) {
            c.interior = kv_info[2 * n + 1] != c.kvl;
            return;
        }
        on_empty(c.item);
        c.item += stride;
    }
}
template <typename F>
__device__ __forceinline__ void a2_advance_lite(A2Lite& c, const A2Shape& sh, const int32_t* __restrict__ kv_info, int stride,
                                                F&& on_empty) {
    if (++c.j == c.nkv) {
        c.j = 0;
        c.item += stride;
        a2_seek_lite(c, sh, kv_info, stride, on_empty);
    }
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
// control thread of one tile: TMA producer + MMA issuer over the tile's stream of KV blocks g = 0, 1, 2, ...
//   s_free(g)  (S(g) is in the softmax registers)  -> issue S(g+1); K stage g & 1 is free -> load K(g+2);
//                                                     block g was its item's last -> its Q buffer is free -> load Q(it+2)
//   p_full(g)  (P(g) is in TMEM)                   -> issue O += P(g) V(g)
//   pv_done(g)                                     -> V stage g & 1 is free -> load V(g+2)
// The softmax threads arrive s_free(g+1) right after p_full(g) (they pull S(g+1) while they exponentiate block g).
// ------------------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void a2_control(uint8_t* slice, const A2Bars& bar, uint32_t tmem_tile, const CUtensorMap* tma_qkv,
                                           int slot, int stride, const A2Shape& sh, const int32_t* __restrict__ kv_info) {
    using Cfg = A2Cfg<D>;
    constexpr uint32_t idesc_s = make_idesc_bf16(A2_BLOCK, A2_BLOCK, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(A2_BLOCK, D, false, true);             // B (= V) is MN-major
    const uint32_t s_q = smem_u32(slice + Cfg::OFF_Q), s_k = smem_u32(slice + Cfg::OFF_K), s_v = smem_u32(slice + Cfg::OFF_V);
    const uint32_t tmem_s = tmem_tile + Cfg::TM_S, tmem_p = tmem_tile + Cfg::TM_P, tmem_o = tmem_tile + Cfg::TM_O;
    const int h = sh.h, k_tokens = sh.k_tokens;
    auto valid = [&](const A2Cur& c) { return a2_valid(c, sh); };
    auto seek = [&](A2Cur& c, int item) {
        const int4 r = a2_seek_ctrl(item, sh.total, sh.nqb, sh.heads, stride, kv_info);
        c.item = r.x; c.n = r.y; c.head = r.z; c.q0 = r.w;
        if (c.item < sh.total) {
            c.kvl = kv_info[2 * c.n];
            c.n_nonpad = kv_info[2 * c.n + 1];
            c.nkv = (c.kvl + A2_BLOCK - 1) / A2_BLOCK;
        }
    };
    auto next_item = [&](A2Cur& c) {
        ++c.it;
        c.j = 0;
        seek(c, c.item + stride);
    };
    auto advance = [&](A2Cur& c) {
        ++c.g;
        if (++c.j == c.nkv) next_item(c);
    };

    auto load_q = [&](const A2Cur& c) {                      // Q of item c.it -> buffer c.it & 1
        const int b = c.it & 1;
        mbar_arrive_expect_tx(&bar.q[b], Cfg::TILE_BYTES);
        tma_load_2d(slice + Cfg::OFF_Q + b * Cfg::TILE_BYTES, tma_qkv, &bar.q[b], c.head * D, c.n * k_tokens + c.q0);
    };
    auto load_k = [&](const A2Cur& c) {                      // K of block c.g -> stage c.g & 1
        const int st = c.g & 1;
        mbar_arrive_expect_tx(&bar.k_full[st], Cfg::TILE_BYTES);
        tma_load_2d(slice + Cfg::OFF_K + st * Cfg::TILE_BYTES, tma_qkv, &bar.k_full[st], h + c.head * D,
                    c.n * k_tokens + c.j * A2_BLOCK);
    };
    auto load_v = [&](const A2Cur& c) {
        const int st = c.g & 1;
        mbar_arrive_expect_tx(&bar.v_full[st], Cfg::TILE_BYTES);
        tma_load_2d(slice + Cfg::OFF_V + st * Cfg::TILE_BYTES, tma_qkv, &bar.v_full[st], 2 * h + c.head * D,
                    c.n * k_tokens + c.j * A2_BLOCK);
    };
    auto issue_s = [&](const A2Cur& c) {                     // S = Q K(g)^T : K-major x K-major, D/16 k-steps
        if (c.j == 0) mbar_wait(&bar.q[c.it & 1], (c.it >> 1) & 1);
        mbar_wait(&bar.k_full[c.g & 1], (c.g >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < D / 16; ++s) {
            const uint64_t qd = make_smem_desc(s_q + (c.it & 1) * Cfg::TILE_BYTES + s * 32, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
            const uint64_t kd = make_smem_desc(s_k + (c.g & 1) * Cfg::TILE_BYTES + s * 32, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
            umma_bf16_ss(tmem_s, qd, kd, idesc_s, s != 0);
        }
        umma_commit(bar.s_full);
    };

    A2Cur cp;                                                // PV cursor = the block the loop is at
    cp.item = slot; cp.it = 0; cp.j = 0; cp.g = 0;
    cp.n = cp.head = cp.q0 = cp.kvl = cp.nkv = cp.n_nonpad = 0;
    seek(cp, slot);
    if (!valid(cp)) return;
    A2Cur cq = cp, ck = cp, cv = cp, cs = cp, cf = cp;       // Q-load (item-granular) / K-load / V-load / S-issue / s_free cursors
    load_q(cq);
    next_item(cq);
    if (valid(cq)) { load_q(cq); next_item(cq); }
    load_k(ck); advance(ck);
    if (valid(ck)) { load_k(ck); advance(ck); }
    load_v(cv); advance(cv);
    if (valid(cv)) { load_v(cv); advance(cv); }
    issue_s(cs); advance(cs);

    auto after_s_free = [&]() {                              // S(cf.g) is in registers
        mbar_wait(bar.s_free, cf.g & 1);
        tc_fence_after();
        if (valid(cs)) { issue_s(cs); advance(cs); }         // S(g+1): runs while the softmax warps exponentiate block g
        if (valid(ck)) { load_k(ck); advance(ck); }          // K(g+2) -> the stage S(g) has been read from
        if (cf.j == cf.nkv - 1 && valid(cq)) {               // the item's last S: its Q buffer takes the item after next
            load_q(cq);
            next_item(cq);
        }
        advance(cf);
    };
    after_s_free();                                          // block 0 (the softmax prologue)
    while (valid(cp)) {
        const int st = cp.g & 1;
        // O += P(g) V(g) : P from tensor memory (lane = query row, 8 packed columns per 16 keys), V MN-major from smem
        mbar_wait(bar.p_full, cp.g & 1);
        mbar_wait(&bar.v_full[st], (cp.g >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < A2_BLOCK / 16; ++s) {
            const uint64_t vd = make_smem_desc(s_v + st * Cfg::TILE_BYTES + s * 16 * Cfg::ROW_BYTES, Cfg::TILE_BYTES,
                                               8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
            umma_bf16_ts(tmem_o, tmem_p + s * 8, vd, idesc_pv, (cp.j | s) != 0);
        }
        umma_commit(&bar.pv_done[st]);
        if (cp.j == cp.nkv - 1) umma_commit(bar.o_full);
        if (valid(cf)) after_s_free();                       // s_free(g+1) follows p_full(g) at once
        mbar_wait(&bar.pv_done[st], (cp.g >> 1) & 1);
        if (valid(cv)) { load_v(cv); advance(cv); }          // V(g+2) -> the stage PV(g) has drained
        advance(cp);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// softmax warps of one tile: thread r owns query row r
// ------------------------------------------------------------------------------------------------------------------
// bit i = key j0 + i of the sequence is not a pad id (interior-pad sequences only: kept out of line, the hot loop must stay
// small enough for the instruction cache -- the first pipelined build inlined it eight times and stalled on instruction fetch)
__device__ __noinline__ uint32_t a2_key_bits(const uint8_t* __restrict__ key_mask, long long row_base, int j0, int k_tokens) {
    uint32_t bits = 0;
    if (((row_base + j0) & 15) == 0 && j0 + 32 <= k_tokens) {
        const uint4* p = reinterpret_cast<const uint4*>(key_mask + row_base + j0);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint4 u = __ldg(p + q);
            const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if ((wd[i >> 2] >> (8 * (i & 3))) & 0xffu) bits |= 1u << (q * 16 + i);
        }
    } else {
        for (int i = 0; i < 32; ++i)
            if (j0 + i < k_tokens && key_mask[row_base + j0 + i] != 0) bits |= 1u << i;
    }
    return bits;
}
// keys of chunk `c` (32 keys) of block `b` that are padding -> -inf
__device__ __forceinline__ void a2_mask_chunk(float* s, int c, const A2Lite& b, const uint8_t* __restrict__ key_mask,
                                              const A2Shape& sh) {
    const int j0 = b.j * A2_BLOCK + c * 32;
    const int lim = b.kvl - j0;                                               // keys of this chunk inside kv_len
    uint32_t valid = lim >= 32 ? 0xffffffffu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
    if (b.interior)                                                           // pad ids before the last real token
        valid &= a2_key_bits(key_mask, static_cast<long long>(b.item / (sh.nqb * sh.heads)) * sh.k_tokens, j0, sh.k_tokens);
    if (valid != 0xffffffffu) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (!((valid >> i) & 1u)) s[i] = -CUDART_INF_F;
    }
}
// running row max over one 32-key chunk: four independent FMNMX3 chains
__device__ __forceinline__ void a2_max_chunk(const float* s, float (&mx4)[4]) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        mx4[0] = a2_max3(mx4[0], s[i], s[i + 1]);
        mx4[1] = a2_max3(mx4[1], s[i + 2], s[i + 3]);
        mx4[2] = a2_max3(mx4[2], s[i + 4], s[i + 5]);
        mx4[3] = a2_max3(mx4[3], s[i + 6], s[i + 7]);
    }
}
// P = exp2(s log2e - m) for one 32-key chunk -> 16 packed bf16 pairs (column c of lane r holds keys 2c, 2c+1 of query row r:
// the K-major A operand of O += P V); the fp32 values are added to the row sum
template <int POLY>
__device__ __forceinline__ void a2_exp_chunk(const float* s, uint64_t sc2, uint64_t nm2, uint64_t (&sum2)[2], uint32_t (&pk)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float x0, x1;
        unpack_f32x2(fma_f32x2(pack_f32x2(s[2 * i], s[2 * i + 1]), sc2, nm2), x0, x1);
        if (i % 4 < POLY) {                                                   // this pair goes to the FMA pipe
            a2_exp2_poly_pair(x0, x1);
        } else {                                                              // this pair goes to the MUFU
            x0 = a2_ex2(x0);
            x1 = a2_ex2(x1);
        }
        sum2[i & 1] = add_f32x2(sum2[i & 1], pack_f32x2(x0, x1));
        pk[i] = pack_bf16x2(x0, x1);
    }
}

template <int D, int POLY>
__device__ __forceinline__ void a2_softmax(const A2Bars& bar, uint32_t tmem_tile, int r, int slot, int stride,
                                           const A2Shape& sh, const int32_t* __restrict__ kv_info,
                                           const uint8_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out,
                                           float* __restrict__ lse2) {
    using Cfg = A2Cfg<D>;
    const uint32_t lane_addr = static_cast<uint32_t>(r & ~31) << 16;          // TMEM lane quarter of this warp
    const uint32_t tmem_s = tmem_tile + Cfg::TM_S + lane_addr, tmem_p = tmem_tile + Cfg::TM_P + lane_addr;
    const uint32_t tmem_o = tmem_tile + Cfg::TM_O + lane_addr;
    const int k_tokens = sh.k_tokens, h = sh.h;
    // (row valid?, output row, log-sum-exp slot) of this thread for a work item: one decode per item
    auto row_of = [&](int item, __nv_bfloat16*& orow, float*& lrow) {
        const int q0 = (item % sh.nqb) * A2_BLOCK, head = (item / sh.nqb) % sh.heads, n = item / (sh.nqb * sh.heads);
        orow = out + (static_cast<long long>(n) * k_tokens + q0 + r) * h + head * D;
        lrow = lse2 == nullptr ? nullptr : lse2 + (static_cast<size_t>(n) * sh.heads + head) * k_tokens + q0 + r;
        return q0 + r < k_tokens;
    };
    auto zero_fill = [&](int item) {                                          // all-pad sequence: the reference never encodes one
        __nv_bfloat16* orow;
        float* lrow;
        if (row_of(item, orow, lrow)) {
            for (int i = 0; i < D / 8; ++i) reinterpret_cast<uint4*>(orow)[i] = make_uint4(0, 0, 0, 0);
            if (lrow != nullptr) *lrow = -CUDART_INF_F;
        }
    };
#ifdef A2_SKEW
    // experiment: start tile 1 A2_SKEW cycles late so that the two softmax warps of a scheduler do not run their exp2 phases
    // in lock-step (both tiles start together and identical work keeps them in phase)
    if (tmem_tile & 256u) {
        const long long t0 = clock64();
        while (clock64() - t0 < A2_SKEW) { }
    }
#endif
    A2Lite c;                                                                 // the block being exponentiated
    c.item = slot; c.j = 0; c.nkv = 0; c.kvl = 0; c.interior = false;
    a2_seek_lite(c, sh, kv_info, stride, zero_fill);
    if (c.item >= sh.total) return;
    int g = 0, it = 0;                                                        // block / item ordinals in the tile's stream

    float a[A2_BLOCK], b[A2_BLOCK];                                           // S / P of the current and of the next block
    float mx;                                                                 // row max of the current block (masked)
    float m_run = -CUDART_INF_F;                                              // running reference max, log2 domain
    float l_run = 0.f;

    // prologue: S of the stream's first block
    mbar_wait(bar.s_full, 0);
    tc_fence_after();
    {
        uint32_t* raw = reinterpret_cast<uint32_t*>(a);
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld32(tmem_s + q * 32, raw + q * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_wait_regs32(raw + q * 32);
        tc_fence_before();
        mbar_arrive(bar.s_free);
        float mx4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            a2_mask_chunk(a + q * 32, q, c, key_mask, sh);
            a2_max_chunk(a + q * 32, mx4);
        }
        mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    }

    // one pipeline step: exponentiate block c (in `cur`, row max `mx`) while pulling the next block into `nxt`
    auto step = [&](float (&cur)[A2_BLOCK], float (&nxt)[A2_BLOCK]) {
        const A2Lite b0 = c;
        A2Lite b1 = c;
        a2_advance_lite(b1, sh, kv_info, stride, zero_fill);
        const bool has_next = b1.item < sh.total;
#ifdef A2_TIMELINE
        const bool tl_on = d_a2_tl != nullptr && r == 0 && blockIdx.x < 8 && g < 64;
        const int tl_base = ((blockIdx.x * 2 + ((tmem_tile >> 8) & 1)) * 64 + g) * 8;
#endif
        TL2(0);
        uint32_t* nraw = reinterpret_cast<uint32_t*>(nxt);
        if (b0.j == 0) { m_run = -CUDART_INF_F; l_run = 0.f; }
        // Lazy rescaling (see attention.cu): the reference max moves only when the row max grew by more than 2^8, so
        // P <= 256 (exact in the fp32 sum, harmless in bf16) and O / l almost never need a correction.
        const float m_cand = fmaxf(m_run, mx * A2_LOG2E);
        float alpha = 1.0f;
        if (m_cand > m_run + 8.0f) {                                          // first valid block: m_run = -inf -> alpha = 0
            alpha = a2_ex2(m_run - m_cand);
            m_run = m_cand;
        }
        const float m_use = (m_run == -CUDART_INF_F) ? 0.f : m_run;
        const uint64_t sc2 = pack_f32x2(A2_LOG2E, A2_LOG2E), nm2 = pack_f32x2(-m_use, -m_use);
        uint64_t sum2[2] = {pack_f32x2(0.f, 0.f), pack_f32x2(0.f, 0.f)};
        float nm4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        uint32_t pk0[16], pk1[16];
        // keys 0..63: exponentials first, so that O += P(g-1) V(g-1) and S(g+1) have time to complete
        a2_exp_chunk<POLY>(cur, sc2, nm2, sum2, pk0);
        a2_exp_chunk<POLY>(cur + 32, sc2, nm2, sum2, pk1);
        TL2(1);
        if (b0.j > 0) {                                                       // PV(g-1) consumed P and finished O
            mbar_wait(&bar.pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);         // (j == 0: the previous item's o_full)
            tc_fence_after();
        }
        TL2(2);
        tmem_st16(tmem_p, pk0);
        tmem_st16(tmem_p + 16, pk1);
        if (has_next) {
            mbar_wait(bar.s_full, (g + 1) & 1);
            tc_fence_after();
        }
        TL2(3);
        tmem_ld32(tmem_s, nraw);                                              // next keys 0..63 -> the registers just vacated
        tmem_ld32(tmem_s + 32, nraw + 32);                                    // (end of stream: reads stale S, never used)
        // keys 64..95
        a2_exp_chunk<POLY>(cur + 64, sc2, nm2, sum2, pk0);
        tmem_st16(tmem_p + 32, pk0);
        tmem_ld_wait_regs32(nraw);
        tmem_ld_wait_regs32(nraw + 32);
        tmem_ld32(tmem_s + 64, nraw + 64);
        TL2(4);
        if (has_next) {
            a2_mask_chunk(nxt, 0, b1, key_mask, sh);
            a2_mask_chunk(nxt + 32, 1, b1, key_mask, sh);
        }
        // keys 96..127, with the row max of the next block's keys 0..63 in its empty issue slots
        a2_max_chunk(nxt, nm4);
        a2_max_chunk(nxt + 32, nm4);
        a2_exp_chunk<POLY>(cur + 96, sc2, nm2, sum2, pk1);
        tmem_st16(tmem_p + 48, pk1);
        tmem_ld_wait_regs32(nraw + 64);
        tmem_ld32(tmem_s + 96, nraw + 96);
        TL2(5);
        if (has_next) a2_mask_chunk(nxt + 64, 2, b1, key_mask, sh);
        a2_max_chunk(nxt + 64, nm4);
        float sa, sb, sc, sd;
        unpack_f32x2(sum2[0], sa, sb);
        unpack_f32x2(sum2[1], sc, sd);
        l_run = l_run * alpha + ((sa + sb) + (sc + sd));
        if (b0.j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {             // rare: rescale the running O accumulator
#pragma unroll
            for (int q = 0; q < D / 16; ++q) {
                uint32_t o[16];
                tmem_ld16(tmem_o + q * 16, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st16(tmem_o + q * 16, o);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar.p_full);                                              // the control thread may issue O += P(g) V(g)
        TL2(6);
        tmem_ld_wait_regs32(nraw + 96);
        if (has_next) {
            tc_fence_before();
            mbar_arrive(bar.s_free);                                          // ... and S(g+2)
            a2_mask_chunk(nxt + 96, 3, b1, key_mask, sh);
        }
        a2_max_chunk(nxt + 96, nm4);
        mx = fmaxf(fmaxf(nm4[0], nm4[1]), fmaxf(nm4[2], nm4[3]));
        if (b0.j == b0.nkv - 1) {
            // item epilogue: O / l -> bf16 -> HBM (packed results live in their own registers, so a store in flight does not
            // hold up the next tcgen05.ld)
            mbar_wait(bar.o_full, it & 1);
            ++it;
            tc_fence_after();
            const float inv_l = 1.0f / l_run;
            __nv_bfloat16* orow;
            float* lrow;
            const bool row_ok = row_of(b0.item, orow, lrow);
            if (row_ok && lrow != nullptr) *lrow = m_run + log2f(l_run);      // row log-sum-exp, log2 domain (backward)
#pragma unroll
            for (int q = 0; q < D / 16; ++q) {                                // 16 columns at a time, alternating registers
                uint32_t o[16];
                tmem_ld16(tmem_o + q * 16, o);
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
                if (row_ok) {
                    reinterpret_cast<uint4*>(orow + q * 16)[0] = u0;
                    reinterpret_cast<uint4*>(orow + q * 16)[1] = u1;
                }
            }
            tc_fence_before();                                                // O is read: the next item's PV may overwrite it
        }
        TL2(7);
        c = b1;
        ++g;
    };

    for (;;) {
        step(a, b);
        if (c.item >= sh.total) break;
        step(b, a);
        if (c.item >= sh.total) break;
    }
}

template <int D, int POLY>
__global__ void __launch_bounds__(A2_THREADS, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tma_qkv, int n_seq, int heads, int k_tokens, int h,
                  const int32_t* __restrict__ kv_info, const uint8_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out,
                  float* __restrict__ lse2) {
    using Cfg = A2Cfg<D>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    A2Shape sh;
    sh.n_seq = n_seq; sh.heads = heads; sh.k_tokens = k_tokens; sh.h = h;
    sh.nqb = (k_tokens + A2_BLOCK - 1) / A2_BLOCK;
    sh.total = n_seq * heads * sh.nqb;

    if (warp == 10) {
        if (lane == 0) {
            if ((smem_u32(smem) & 1023u) != 0) { printf("molly attention2: smem base not 1024-B aligned\n"); __trap(); }
            tma_prefetch_desc(&tma_qkv);
            for (int t = 0; t < 2; ++t) {
                const A2Bars b = a2_bars(bars + t * Cfg::NBAR);
                for (int i = 0; i < 2; ++i) {
                    mbar_init(&b.q[i], 1);
                    mbar_init(&b.k_full[i], 1);
                    mbar_init(&b.v_full[i], 1);
                    mbar_init(&b.pv_done[i], 1);
                }
                mbar_init(b.s_full, 1);
                mbar_init(b.p_full, 128);
                mbar_init(b.o_full, 1);
                mbar_init(b.s_free, 128);
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

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A2_REGS_CONTROL));
        if (warp < 10 && lane == 0) {
            const int t = warp - 8;
            a2_control<D>(smem + t * Cfg::SLICE_BYTES, a2_bars(bars + t * Cfg::NBAR), tmem_base + t * Cfg::TM_TILE, &tma_qkv,
                          2 * blockIdx.x + t, stride, sh, kv_info);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(A2_REGS_SOFTMAX));
        const int t = warp >> 2;
        a2_softmax<D, POLY>(a2_bars(bars + t * Cfg::NBAR), tmem_base + t * Cfg::TM_TILE, threadIdx.x & 127,
                            2 * blockIdx.x + t, stride, sh, kv_info, key_mask, out, lse2);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 10) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int a2_env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e == nullptr ? dflt : atoi(e);
}

template <int D>
int launch_attention2(const CUtensorMap& tm, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                      const uint8_t* key_mask, void* out, float* lse, cudaStream_t stream) {
    using Cfg = A2Cfg<D>;
    static int poly = -1;             // pairs out of 4 whose exp2 runs on the FMA pipe (MOLLY_ATTN_POLY = 0 | 1 | 2)
    if (poly < 0) {
        poly = a2_env_int("MOLLY_ATTN_POLY", ATTENTION2_POLY_DEFAULT);
        if (poly < 0 || poly > 2) poly = ATTENTION2_POLY_DEFAULT;
    }
    auto kernel = poly == 0 ? attention2_kernel<D, 0> : (poly == 1 ? attention2_kernel<D, 1> : attention2_kernel<D, 2>);
    static bool configured = false;
    if (!configured) {
        MOLLY_CUDA(cudaFuncSetAttribute(attention2_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention2_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention2_kernel<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
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
        case 16: return launch_attention2<16>(maps.q, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        case 32: return launch_attention2<32>(maps.q, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        case 64: return launch_attention2<64>(maps.q, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        default: MOLLY_CHECK(false, MOLLY_ERR_UNSUPPORTED, "attention2: head_dim %d unsupported", h / heads);
    }
}

}  // namespace molly
