// Fused bidirectional multi-head attention over K-token padded sequences with per-sequence key masking
// (HF:257-282 eager attention with the additive key-padding mask of HF:679-709), flash-style online softmax.
//
// Input is the packed bf16 QKV activation [n_seq*k_tokens, 3h] written by the fused QKV GEMM (q already scaled by
// d^-1/2 through the packed weights, rotary already applied); output is bf16 [n_seq*k_tokens, h].
//
// attention_kernel: a CTA streams work items = 128 query rows of one (sequence, head); two CTAs per SM:
//   * TMA brings Q once per item and K/V tiles (128 keys) through a 2-stage mbarrier ring that runs across items
//   * S = Q K^T  : tcgen05.mma, both operands K-major from swizzled smem, fp32 accumulator in TMEM (128 columns)
//   * softmax    : 4 warps, thread r owns row r (tcgen05.ld 32x32b: TMEM lane == row, so row max / row sum need no
//                  shuffles); exp2 with the running max in the log2 domain; P written to smem as the bf16 K-major
//                  A operand (128-B swizzle); O rescaled in TMEM only when a warp's running max moved
//   * O += P V   : tcgen05.mma, B operand = V tile straight from TMA ([key][d] row-major == MN-major B), fp32 in TMEM
//   * keys >= kv_len are never loaded (whole blocks skipped); interior pad keys use the byte mask
// Two CTAs are co-resident per SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
// Variants kept for A/B timing (tests/test_gpu_variants.py): 64-key blocks / three CTAs per SM (MOLLY_ATTN_KVB=64), one CTA per
// item (MOLLY_ATTN_STREAM=0), exp2 partly on the FMA pipe (MOLLY_ATTN_POLY), register-pipelined 64-key kernel (MOLLY_ATTN_PIPE=1).
#include <math_constants.h>
#include <stdlib.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace molly {

namespace {

// Bring-up aid: -DATT_ABLATE=n builds a deliberately WRONG kernel with one piece removed, to time what that piece costs
// (tools/attn_ablate.sh), a bit mask: 1 no exp2, 2 no P pack/store, 4 PV MMA one k-step, 8 no row max, 16 S MMA one k-step,
// 32 no K/V TMA after the first two blocks.
#ifndef ATT_ABLATE
#define ATT_ABLATE 0
#endif

// -DATT_TIMELINE: clock64 timeline of the first 2048 CTAs of the default kernel (tools/attn_timeline.py), 72 slots each.
#ifdef ATT_TIMELINE
__device__ long long* d_attn_tl = nullptr;
#define TL(slot)                                                                                               \
    do {                                                                                                       \
        if (d_attn_tl != nullptr && tl_id < 2048 && (slot) < 72) d_attn_tl[tl_id * 72 + (slot)] = clock64();   \
    } while (0)
#else
#define TL(slot) do { } while (0)
#endif

// A/B build switches (tools/build_variant.sh): -DATT_DIRECT_EPI=1 stores O rows straight from registers (round-1 epilogue),
// -DATT_MAX2=1 two-input row max, -DATT_THREAD_ARRIVE=1 every softmax thread arrives on s_free / p_full (128 arrivals)
#ifndef ATT_DIRECT_EPI
#define ATT_DIRECT_EPI 0
#endif
#ifndef ATT_MAX2
#define ATT_MAX2 0
#endif
#ifndef ATT_THREAD_ARRIVE
#ifndef ATT_TS
#define ATT_TS 0        // 1: P (bf16) lives in the tensor-memory columns S and O leave free and feeds O += P V as the A operand
#endif                  //    from tensor memory (TS-form MMA) when they fit: KVB + D + KVB/2 <= 256 (head_dim <= 64 at 128 keys).
                        //    Measured 555 us against 542 (it costs 80 B of spills at the 168-register cap): off.  The same change
                        //    is worth 11-13 % in the attention BACKWARD kernels, whose small MMAs are bound by operand reads.
#ifndef ATT_ELECT
#define ATT_ELECT 1
#endif
#define ATT_THREAD_ARRIVE 1       // measured (profiles/r02_attention.md): one elected arrival per warp is SLOWER with the staged epilogue
#endif
constexpr int ATT_ARRIVALS = ATT_THREAD_ARRIVE ? 128 : 4;       // arrivals per phase of s_free / p_full

constexpr int ATT_BLOCK = 128;                 // query rows per CTA == keys per KV block
constexpr int ATT_THREADS = 160;               // 4 softmax warps + 1 control warp
constexpr float LOG2E = 1.4426950408889634f;
constexpr int ATTN_POLY_DEFAULT = 0;           // set from the A/B measurement (tools/attn_bench.py)
constexpr bool ATTN_PIPE_DEFAULT = false;      // register-pipelined 64-key kernel at head_dim <= 64 (set from the A/B measurement)
constexpr bool ATTN_KVB64_DEFAULT = false;     // 64-key KV blocks (3 CTAs / SM) at head_dim <= 64

template <int D, int KVB = 128>       // KVB = keys per KV block (128, or 64 to fit three CTAs per SM at head_dim <= 64)
struct AttnCfg {
    static constexpr int BOX_D = D < 64 ? D : 64;
    static constexpr int NBOX = D / BOX_D;
    static constexpr int ROW_BYTES = BOX_D * 2;                       // 32 / 64 / 128
    static constexpr uint32_t LAYOUT = ROW_BYTES == 128 ? kLayoutSW128 : (ROW_BYTES == 64 ? kLayoutSW64 : kLayoutSW32);
    static constexpr int BOX_BYTES = ATT_BLOCK * ROW_BYTES;           // one TMA box of the Q tile
    static constexpr int TILE_BYTES = NBOX * BOX_BYTES;               // the Q tile
    static constexpr int KV_BOX_BYTES = KVB * ROW_BYTES;
    static constexpr int KV_TILE_BYTES = NBOX * KV_BOX_BYTES;         // one K / V tile
    static constexpr int P_BYTES = ATT_BLOCK * KVB * 2;               // KVB/64 bf16 swizzle atoms of 128 x 64
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = TILE_BYTES;
    static constexpr int OFF_V = OFF_K + 2 * KV_TILE_BYTES;
    static constexpr int OFF_P = OFF_V + 2 * KV_TILE_BYTES;
    static constexpr int OFF_BAR = OFF_P + P_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 128;
    static constexpr int TMEM_COLS = (KVB + D) <= 128 ? 128 : 256;    // S: KVB columns, O: D columns
    static constexpr int BY_SMEM = (227 * 1024) / (SMEM_BYTES + 1024);
    static constexpr int BY_TMEM = 512 / TMEM_COLS;
    static constexpr int MIN_CTAS_RAW = BY_SMEM < BY_TMEM ? BY_SMEM : BY_TMEM;
    static constexpr int MIN_CTAS = MIN_CTAS_RAW > 3 ? 3 : (MIN_CTAS_RAW < 1 ? 1 : MIN_CTAS_RAW);
};

// Hand-off of a softmax warp: every lane has fenced its tcgen05 / shared-memory work; ONE lane arrives for the warp
// (128 single-thread arrivals are 128 serialized updates of the same mbarrier word before the control thread wakes up).
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
#if ATT_THREAD_ARRIVE
    mbar_arrive(bar);
#else
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
#endif
}

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float max3(float a, float b, float c) {             // FMNMX3
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
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

// O / l -> bf16 -> HBM for one warp's 32 query rows.  A thread owns a whole row in TMEM, and storing it directly costs 32
// scattered 16-B pieces per STG (one per row, row pitch h * 2 bytes): the round-2 timeline showed ~4 000 cycles per work item
// in this epilogue.  Instead the rows are staged in shared memory -- the warp's OWN 4 KB slice of the P buffer(s), free
// between the item's last O += P V and the next item's first P store, touched by no other warp -- and written out with
// D/8 lanes per row, so every STG covers whole 128-B lines.  16-B chunks are XOR-swizzled by the row: conflict-free.
template <int D, int KVB>
__device__ __forceinline__ void store_o_rows(uint8_t* p_buf, uint32_t tmem_o_lane, float inv_l, int warp, int lane,
                                             __nv_bfloat16* __restrict__ out_tile, int rows_valid, int h) {
    constexpr int CPR = D / 8;                                 // 16-B chunks per row
    static_assert(D <= 64 || KVB == 128, "head_dim 128 stages through both 16 KB atoms of the P buffer");
    uint8_t* base = p_buf + warp * 4096;
    auto chunk_addr = [&](int lr, int c) -> uint8_t* {
        if constexpr (D >= 64) return base + (c >> 3) * (ATT_BLOCK * 128) + lr * 128 + (((c & 7) ^ (lr & 7)) << 4);
        else return base + lr * (CPR * 16) + (c << 4);
    };
#pragma unroll
    for (int c = 0; c < D / 16; ++c) {
        uint32_t o[16];
        tmem_ld16(tmem_o_lane + c * 16, o);
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
        *reinterpret_cast<uint4*>(chunk_addr(lane, 2 * c)) = u0;
        *reinterpret_cast<uint4*>(chunk_addr(lane, 2 * c + 1)) = u1;
    }
    __syncwarp();
    constexpr int RPI = 32 / CPR;                              // rows per store instruction
#pragma unroll
    for (int k = 0; k < CPR; ++k) {
        const int lr = k * RPI + lane / CPR, c = lane % CPR;
        const uint4 v = *reinterpret_cast<const uint4*>(chunk_addr(lr, c));
        if (warp * 32 + lr < rows_valid)
            *reinterpret_cast<uint4*>(out_tile + static_cast<size_t>(warp * 32 + lr) * h + c * 8) = v;
    }
    __syncwarp();                                              // the slice goes back to P
}

// One KV block of the online softmax for the 128 threads of a softmax group (thread = query row): S(g) from TMEM ->
// mask -> running max (lazy rescale) -> P = exp2(S*log2e - m) as the bf16 K-major A operand in smem -> hand-off.
template <int D, int KVB, int POLY>
__device__ __forceinline__ void softmax_block(uint64_t* bar_s_full, uint32_t s_parity, uint64_t* bar_s_free,
                                              uint64_t* bar_pv_done, uint32_t pv_parity, uint64_t* bar_p_full,
                                              uint32_t tmem_s, uint32_t tmem_o, uint32_t lane_addr, uint8_t* p_row,
                                              int sw, bool interior, const uint8_t* __restrict__ key_mask,
                                              long long row_base, int j0, int kvl, int k_tokens, bool first,
                                              float& m_run, float& l_run, bool tl_on, int tl_id, int tl_base) {
    (void)tl_on; (void)tl_id; (void)tl_base;
    mbar_wait(bar_s_full, s_parity);
    if (tl_on) TL(tl_base + 0);
    tc_fence_after();
    float s[KVB];
    {
        uint32_t raw[KVB];
#pragma unroll
        for (int c = 0; c < KVB / 32; ++c) tmem_ld32(tmem_s + lane_addr + c * 32, raw + c * 32);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < KVB; ++i) s[i] = __uint_as_float(raw[i]);
    }
    tc_fence_before();
    warp_arrive(bar_s_free);                         // S(j) is in registers: the MMA warp may start S(j+1)
    if (tl_on) TL(tl_base + 1);
    if (interior) {
        const uint4* mk = reinterpret_cast<const uint4*>(key_mask + row_base + j0);
        const bool vec_ok = ((row_base + j0) & 15) == 0 && j0 + KVB <= k_tokens;
#pragma unroll
        for (int g = 0; g < KVB / 16; ++g) {
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
    } else if (j0 + KVB > kvl) {
        const int lim = kvl - j0;
#pragma unroll
        for (int i = 0; i < KVB; ++i)
            if (i >= lim) s[i] = -CUDART_INF_F;
    }
    float mx4[4] = {s[0], s[1], s[2], s[3]};          // 4 independent chains of 3-input max (FMNMX3)
#if ATT_MAX2
#pragma unroll
    for (int i = 4; i < KVB; i += 4) {
        mx4[0] = fmaxf(mx4[0], s[i]); mx4[1] = fmaxf(mx4[1], s[i + 1]);
        mx4[2] = fmaxf(mx4[2], s[i + 2]); mx4[3] = fmaxf(mx4[3], s[i + 3]);
    }
#else
#pragma unroll
    for (int i = 4; i + 8 <= KVB; i += 8) {
        mx4[0] = max3(mx4[0], s[i], s[i + 1]); mx4[1] = max3(mx4[1], s[i + 2], s[i + 3]);
        mx4[2] = max3(mx4[2], s[i + 4], s[i + 5]); mx4[3] = max3(mx4[3], s[i + 6], s[i + 7]);
    }
    mx4[0] = max3(mx4[0], s[KVB - 4], s[KVB - 3]);
    mx4[1] = max3(mx4[1], s[KVB - 2], s[KVB - 1]);
#endif
#if ATT_ABLATE & 8
    const float mx = fmaxf(s[0], s[KVB - 1]);
#else
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
#endif
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
    if (tl_on) TL(tl_base + 2);
    constexpr bool TS = ATT_TS && (KVB + D + KVB / 2 <= AttnCfg<D, KVB>::TMEM_COLS);
    uint32_t pk[TS ? KVB / 2 : 1];                    // TS: P is packed to bf16 pairs as it is produced (half the live registers)
    const uint64_t sc2 = pack_f32x2(LOG2E, LOG2E), nm2 = pack_f32x2(-m_use, -m_use);
    uint64_t sum2[2] = {pack_f32x2(0.f, 0.f), pack_f32x2(0.f, 0.f)};
#pragma unroll
    for (int i = 0; i < KVB; i += 4) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float x0, x1;
            unpack_f32x2(fma_f32x2(pack_f32x2(s[i + 2 * u], s[i + 2 * u + 1]), sc2, nm2), x0, x1);
            if (((i >> 1) + u) % 4 < POLY) {          // this pair goes to the FMA pipe
                exp2_poly_pair(x0, x1);
                s[i + 2 * u] = x0;
                s[i + 2 * u + 1] = x1;
            } else {                                  // this pair goes to the MUFU
#if ATT_ABLATE & 1
                s[i + 2 * u] = x0;
                s[i + 2 * u + 1] = x1;
#else
                s[i + 2 * u] = ex2(x0);
                s[i + 2 * u + 1] = ex2(x1);
#endif
            }
            sum2[u] = add_f32x2(sum2[u], pack_f32x2(s[i + 2 * u], s[i + 2 * u + 1]));
            if constexpr (TS) pk[(i >> 1) + u] = pack_bf16x2(s[i + 2 * u], s[i + 2 * u + 1]);
        }
    }
    float sa, sb, sc, sd;
    unpack_f32x2(sum2[0], sa, sb);
    unpack_f32x2(sum2[1], sc, sd);
    l_run = l_run * alpha + ((sa + sb) + (sc + sd));
    if (tl_on) TL(tl_base + 3);
    // PV(j-1) must have drained the P buffer and finished O before either is touched again
    if (!first) {                                     // (first block: covered by the previous item's bar_o_full)
        mbar_wait(bar_pv_done, pv_parity);
        tc_fence_after();
    }
    if constexpr (TS) {
        // P -> tensor memory, bf16, two keys per 32-bit column, right behind O
        const uint32_t tmem_p = tmem_o + D;
#pragma unroll
        for (int c = 0; c < KVB / 16; ++c) tmem_st8(tmem_p + lane_addr + c * 8, pk + c * 8);
    } else {
    // P -> smem, bf16, K-major with the 128-B swizzle: 16-B chunk c of row r lands at chunk (c ^ (r & 7))
#if ATT_ABLATE & 2
    if (l_run == 123.456f)
#endif
#pragma unroll
    for (int a = 0; a < KVB / 64; ++a) {
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
    }
    // rescale the running O accumulator (complete through block j-1, see the wait above)
    if (!first && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
        for (int c = 0; c < D / 16; ++c) {
            uint32_t o[16];
            tmem_ld16(tmem_o + lane_addr + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_o + lane_addr + c * 16, o);
        }
        if constexpr (!TS) tmem_st_wait();
    }
    if constexpr (TS) tmem_st_wait(); else fence_proxy_async_smem();
    tc_fence_before();
    warp_arrive(bar_p_full);
    if (tl_on) TL(tl_base + 4);
}

// LSE: also write the row log-sum-exp (training forward).  A template flag, not a runtime test: the inference instantiation must
// stay exactly at the 168-register cap of two CTAs per SM (one more live pointer makes it spill: -9 % measured).
template <int D, int KVB, int POLY, bool LSE = false>
__global__ void __launch_bounds__(ATT_THREADS, AttnCfg<D, KVB>::MIN_CTAS)
attention_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_kv, int n_seq,
                 int heads, int k_tokens, int h, const int32_t* __restrict__ kv_info,
                 const uint8_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out, float* __restrict__ lse2) {
    using Cfg = AttnCfg<D, KVB>;
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
    // Work items = (sequence, head, 128-query block), query block fastest so that the CTAs running side by side share one
    // (sequence, head)'s K/V in L2.  A CTA walks items blockIdx.x, blockIdx.x + gridDim.x, ...; all their KV blocks form ONE
    // stream g = 0, 1, 2, ... that indexes the K/V ring and every barrier parity, so the next item's Q / K / V loads and its
    // first S = QK^T are in flight while the softmax warps still normalise and store the current item's O.
    const int nqb = (k_tokens + ATT_BLOCK - 1) / ATT_BLOCK;
    const int total = n_seq * heads * nqb;
    struct Item { int n, head, q0, kvl, nkv, n_nonpad; };
    auto decode = [&](int item) {
        Item w;
        w.q0 = (item % nqb) * ATT_BLOCK;
        w.head = (item / nqb) % heads;
        w.n = item / (nqb * heads);
        w.kvl = kv_info[2 * w.n];
        w.n_nonpad = kv_info[2 * w.n + 1];
        w.nkv = (w.kvl + KVB - 1) / KVB;
        return w;
    };
    int tl_id = blockIdx.x;
    (void)tl_id;
    if (threadIdx.x == 0) TL(0);

    if (warp == 4) {
        if (lane == 0) {
            if ((smem_u32(smem) & 1023u) != 0) { printf("molly attention: smem base not 1024-B aligned\n"); __trap(); }
            tma_prefetch_desc(&tma_qkv);
            tma_prefetch_desc(&tma_kv);
            mbar_init(bar_q, 1);
            mbar_init(&bar_kv_full[0], 1); mbar_init(&bar_kv_full[1], 1);
            mbar_init(&bar_kv_empty[0], 1); mbar_init(&bar_kv_empty[1], 1);
            mbar_init(bar_s_full, 1);
            mbar_init(bar_p_full, ATT_ARRIVALS);
            mbar_init(bar_o_full, 1);
            mbar_init(bar_s_free, ATT_ARRIVALS);
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
    if (threadIdx.x == 0) TL(1);
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + KVB;

    if (warp == 4) {
        // elect.sync, not `lane == 0`: ptxas then knows ONE thread runs this region and issues the tcgen05 / TMA instructions
        // back to back; under a lane test it wraps EACH of them in an elect / branch loop over the "active" threads (13
        // instructions and a taken branch per MMA -- the control thread, not the tensor pipe, paced the small MMAs)
        if (ATT_ELECT ? elect_one() : lane == 0) {
            // ---------------- control thread: TMA producer + MMA issuer ----------------
            constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BLOCK, KVB, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BLOCK, D, false, true);      // B (= V) is MN-major
            const uint32_t s_q = smem_u32(smem + Cfg::OFF_Q), s_k = smem_u32(smem + Cfg::OFF_K);
            const uint32_t s_v = smem_u32(smem + Cfg::OFF_V), s_p = smem_u32(smem + Cfg::OFF_P);
            auto next_item = [&](int item) {                 // first item >= `item` of this CTA that has keys
                while (item < total && kv_info[2 * (item / (nqb * heads))] <= 0) item += gridDim.x;
                return item;
            };
            auto load_q = [&](const Item& w) {               // Q: 128-row boxes
                mbar_arrive_expect_tx(bar_q, Cfg::TILE_BYTES);
#pragma unroll
                for (int b = 0; b < Cfg::NBOX; ++b)
                    tma_load_2d(smem + Cfg::OFF_Q + b * Cfg::BOX_BYTES, &tma_qkv, bar_q, w.head * D + b * Cfg::BOX_D,
                                w.n * k_tokens + w.q0);
            };
            // load cursor: runs two KV blocks ahead of the compute cursor, across item boundaries
            int l_item = next_item(blockIdx.x), l_j = 0, g_load = 0;
            Item lw = decode(l_item < total ? l_item : 0);
            auto load_next_kv = [&]() {                      // K / V: KVB-row boxes into stage g_load & 1
                if (l_item >= total) return;
                if ((ATT_ABLATE & 32) && g_load >= 2) {
                    mbar_arrive(&bar_kv_full[g_load & 1]);
                } else {
                    const int stg = g_load & 1, row = lw.n * k_tokens + l_j * KVB;
                    mbar_arrive_expect_tx(&bar_kv_full[stg], 2 * Cfg::KV_TILE_BYTES);
#pragma unroll
                    for (int b = 0; b < Cfg::NBOX; ++b) {
                        tma_load_2d(smem + Cfg::OFF_K + stg * Cfg::KV_TILE_BYTES + b * Cfg::KV_BOX_BYTES, &tma_kv,
                                    &bar_kv_full[stg], h + lw.head * D + b * Cfg::BOX_D, row);
                        tma_load_2d(smem + Cfg::OFF_V + stg * Cfg::KV_TILE_BYTES + b * Cfg::KV_BOX_BYTES, &tma_kv,
                                    &bar_kv_full[stg], 2 * h + lw.head * D + b * Cfg::BOX_D, row);
                    }
                }
                ++g_load;
                if (++l_j == lw.nkv) {
                    l_item = next_item(l_item + gridDim.x);
                    l_j = 0;
                    if (l_item < total) lw = decode(l_item);
                }
            };
            auto issue_s = [&](int g) {          // S = Q K(g)^T : K-major x K-major, D/16 k-steps
                mbar_wait(&bar_kv_full[g & 1], (g >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < ((ATT_ABLATE & 16) ? 1 : D / 16); ++s) {
                    const uint32_t off = ((s * 16) / Cfg::BOX_D) * Cfg::BOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    const uint32_t koff = ((s * 16) / Cfg::BOX_D) * Cfg::KV_BOX_BYTES + ((s * 16) % Cfg::BOX_D) * 2;
                    const uint64_t qd = make_smem_desc(s_q + off, 16, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
                    const uint64_t kd = make_smem_desc(s_k + (g & 1) * Cfg::KV_TILE_BYTES + koff, 16, 8 * Cfg::ROW_BYTES,
                                                       Cfg::LAYOUT);
                    umma_bf16_ss(tmem_s, qd, kd, idesc_s, s != 0);
                }
                umma_commit(bar_s_full);
            };
            int c_item = l_item, it = 0, g = 0;
            if (c_item < total) {
                load_q(lw);
                load_next_kv();
                load_next_kv();
                TL(40);
                mbar_wait(bar_q, 0);
                TL(41);
                issue_s(0);
                TL(42);
            }
            while (c_item < total) {
                const Item w = decode(c_item);
                const int nxt = next_item(c_item + gridDim.x);
                tl_id = c_item;
                for (int j = 0; j < w.nkv; ++j, ++g) {
                    const int st = g & 1;
                    const bool last = j == w.nkv - 1;
                    // (1) the moment the softmax warps hold S(g) in registers, S(g+1) is issued: it runs under softmax(g).
                    //     On an item's last block Q is dead instead: the next item's Q is fetched into the same buffer.
                    mbar_wait(bar_s_free, g & 1);
                    TL(43 + 3 * j);
                    tc_fence_after();
                    if (!last) issue_s(g + 1);
                    else if (nxt < total) load_q(decode(nxt));
                    // (2) O += P(g) V(g) : P K-major (64-key atoms), V MN-major; KVB/16 k-steps of 16 keys
                    mbar_wait(bar_p_full, g & 1);
                    TL(44 + 3 * j);
                    tc_fence_after();
#pragma unroll
                    for (int s = 0; s < ((ATT_ABLATE & 4) ? 1 : KVB / 16); ++s) {
                        const uint64_t vd = make_smem_desc(s_v + st * Cfg::KV_TILE_BYTES + s * 16 * Cfg::ROW_BYTES,
                                                           Cfg::KV_BOX_BYTES, 8 * Cfg::ROW_BYTES, Cfg::LAYOUT);
                        if constexpr (ATT_TS && (KVB + D + KVB / 2 <= Cfg::TMEM_COLS)) {   // A = 16 keys = 8 packed TMEM columns
                            umma_bf16_ts(tmem_o, tmem_o + D + s * 8, vd, idesc_pv, (j | s) != 0);
                        } else {
                            const uint64_t pd = make_smem_desc(s_p + (s >> 2) * (ATT_BLOCK * 128) + (s & 3) * 32, 16, 1024,
                                                               kLayoutSW128);
                            umma_bf16_ss(tmem_o, pd, vd, idesc_pv, (j | s) != 0);
                        }
                    }
                    umma_commit(&bar_kv_empty[st]);          // PV(g) done: stage st, the P buffer and O are free
                    if (last) {
                        umma_commit(bar_o_full);
                        if (nxt < total) {                   // first S of the next item, under this item's epilogue
                            mbar_wait(bar_q, (it + 1) & 1);
                            issue_s(g + 1);
                        }
                    }
                    // (3) refill stage st with stream block g+2 once PV(g) has drained it
                    if (l_item < total) {
                        mbar_wait(&bar_kv_empty[st], (g >> 1) & 1);
                        TL(45 + 3 * j);
                        load_next_kv();
                    }
                }
                ++it;
                c_item = nxt;
            }
        }
    } else {
        // ---------------- softmax warps: thread r owns query row r ----------------
        const int r = threadIdx.x;
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        uint8_t* p_row = smem + Cfg::OFF_P + (r >> 3) * 1024 + (r & 7) * 128;
        const int sw = r & 7;
        int it = 0, g = 0;
#ifdef ATT_SKEW
        // experiment: the second resident CTA of every SM starts ATT_SKEW cycles late (exp2 phases out of lock-step)
        if (2 * blockIdx.x >= gridDim.x) {
            const long long t0 = clock64();
            while (clock64() - t0 < ATT_SKEW) { }
        }
#endif
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const Item w = decode(item);
        const int kvl = w.kvl, nkv = w.nkv, q0 = w.q0, head = w.head;
        const bool interior = w.n_nonpad != w.kvl;            // pad ids before the last real token
        const long long row_base = static_cast<long long>(w.n) * k_tokens;
        tl_id = item;
        if (nkv == 0) {                                       // all-pad sequence: the reference never encodes one
            if (q0 + r < k_tokens) {
                uint4* o = reinterpret_cast<uint4*>(out + (row_base + q0 + r) * h + head * D);
                for (int i = 0; i < D / 8; ++i) o[i] = make_uint4(0, 0, 0, 0);
                if constexpr (LSE) lse2[(static_cast<size_t>(w.n) * heads + head) * k_tokens + q0 + r] = -CUDART_INF_F;
            }
            continue;
        }
        float m_run = -CUDART_INF_F;      // running max, log2 domain
        float l_run = 0.f;
        for (int j = 0; j < nkv; ++j, ++g)
            softmax_block<D, KVB, POLY>(bar_s_full, g & 1, bar_s_free, &bar_kv_empty[(g - 1) & 1],
                                               ((g - 1) >> 1) & 1, bar_p_full, tmem_s, tmem_o, lane_addr, p_row, sw,
                                               interior, key_mask, row_base, j * KVB, kvl, k_tokens, j == 0, m_run, l_run,
                                        r == 0 && j < 6, tl_id, 2 + 5 * j);
        // epilogue: O / l -> bf16 -> HBM
        mbar_wait(bar_o_full, it & 1);
        ++it;
        if (r == 0) TL(34);
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        // row log-sum-exp in the log2 domain (P = exp2(S log2e - lse2)): what the attention backward needs
        if constexpr (LSE) {
            if (q0 + r < k_tokens) lse2[(static_cast<size_t>(w.n) * heads + head) * k_tokens + q0 + r] = m_run + log2f(l_run);
        }
#if ATT_DIRECT_EPI
        {
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
#else
        store_o_rows<D, KVB>(smem + Cfg::OFF_P, tmem_o + lane_addr, inv_l, warp, lane,
                             out + (row_base + q0) * h + head * D, k_tokens - q0, h);
#endif
        tc_fence_before();                                    // O is read: the next item's PV may overwrite it
        }   // item loop
    }

    if (threadIdx.x == 0) TL(35);
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
#ifdef ATT_TIMELINE
    if (threadIdx.x == 0) {
        TL(36);
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (d_attn_tl != nullptr && blockIdx.x < 2048) d_attn_tl[blockIdx.x * 72 + 37] = smid;
    }
#endif
}

// ================================================================================================
// Register-pipelined kernel (head_dim <= 64, 64-key blocks).  The bound of attention_kernel is not a pipe but the serial
// chain of ONE softmax warp per scheduler: wait S -> tcgen05.ld (~300 cycles) -> row max -> exp2 -> pack/store P -> fence ->
// arrive (tools/attn_ablate.sh, -DATT_TIMELINE).  Here a thread keeps TWO 64-key blocks of its row in registers: while the
// exp2 phase of block g runs, the tcgen05.ld of S(g+1) is in flight, and its mask + row max are taken off the chain too
// (they run after P(g) is handed over, while the tensor core does O += P(g) V(g)).  What stays serial per 64 keys is
// exp2 + pack/store + hand-off.  The control thread issues S one block ahead of the PV it is waiting for
// (arrival order: s_free(g+1), p_full(g), s_free(g+2), ...), over a 4-stage K/V ring that runs across work items.
// TMEM: S [0,64) | O [64, 64+D).  Two CTAs per SM (registers), 96 KB smem each.
// ================================================================================================
constexpr int PIPE_KVB = 64;
constexpr int PIPE_NST = 4;

template <int D>
struct AttnPipeCfg {
    using B = AttnCfg<D, PIPE_KVB>;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = B::TILE_BYTES;
    static constexpr int OFF_V = OFF_K + PIPE_NST * B::KV_TILE_BYTES;
    static constexpr int OFF_P = OFF_V + PIPE_NST * B::KV_TILE_BYTES;
    static constexpr int OFF_BAR = OFF_P + B::P_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static constexpr int TMEM_COLS = 128;
};

// mask the keys of one 64-key block (suffix padding by kv_len, interior pad ids by the byte mask) and return the row max
__device__ __forceinline__ float pipe_mask_max(float (&s)[PIPE_KVB], bool interior, const uint8_t* __restrict__ key_mask,
                                               long long row_base, int j0, int kvl, int k_tokens) {
    if (interior) {
        const uint4* mk = reinterpret_cast<const uint4*>(key_mask + row_base + j0);
        const bool vec_ok = ((row_base + j0) & 15) == 0 && j0 + PIPE_KVB <= k_tokens;
#pragma unroll
        for (int g = 0; g < PIPE_KVB / 16; ++g) {
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
    } else if (j0 + PIPE_KVB > kvl) {
        const int lim = kvl - j0;
#pragma unroll
        for (int i = 0; i < PIPE_KVB; ++i)
            if (i >= lim) s[i] = -CUDART_INF_F;
    }
    float mx4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
    for (int i = 4; i < PIPE_KVB; i += 4) {
        mx4[0] = fmaxf(mx4[0], s[i]); mx4[1] = fmaxf(mx4[1], s[i + 1]);
        mx4[2] = fmaxf(mx4[2], s[i + 2]); mx4[3] = fmaxf(mx4[3], s[i + 3]);
    }
    return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
}

// lazy rescaling (see softmax_block): returns alpha for this block and moves m_run only when the row max grew by > 2^8
__device__ __forceinline__ float pipe_update_max(float mx, float& m_run) {
    const float m_cand = fmaxf(m_run, mx * LOG2E);
    float alpha = 1.0f;
    if (m_cand > m_run + 8.0f) {
        alpha = ex2(m_run - m_cand);
        m_run = m_cand;
    }
    return alpha;
}

template <int D, int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_pipe_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_kv, int n_seq,
                      int heads, int k_tokens, int h, const int32_t* __restrict__ kv_info,
                      const uint8_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out) {
    using Cfg = AttnPipeCfg<D>;
    using B = typename Cfg::B;
    constexpr int KVB = PIPE_KVB, NST = PIPE_NST;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* bar_q = bars + 0;
    uint64_t* bar_kv_full = bars + 1;               // [NST]
    uint64_t* bar_kv_empty = bars + 1 + NST;        // [NST]  PV(g) complete: stage g % NST, the P buffer and O are free
    uint64_t* bar_s_full = bars + 1 + 2 * NST;
    uint64_t* bar_s_free = bars + 2 + 2 * NST;      // 128 arrivals: S(g) is in registers
    uint64_t* bar_p_full = bars + 3 + 2 * NST;      // 128 arrivals
    uint64_t* bar_o_full = bars + 4 + 2 * NST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * NST);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqb = (k_tokens + ATT_BLOCK - 1) / ATT_BLOCK;
    const int total = n_seq * heads * nqb;
    struct Item { int n, head, q0, kvl, nkv, n_nonpad; };
    auto decode = [&](int item) {
        Item w;
        w.q0 = (item % nqb) * ATT_BLOCK;
        w.head = (item / nqb) % heads;
        w.n = item / (nqb * heads);
        w.kvl = kv_info[2 * w.n];
        w.n_nonpad = kv_info[2 * w.n + 1];
        w.nkv = (w.kvl + KVB - 1) / KVB;
        return w;
    };

    if (warp == 4) {
        if (lane == 0) {
            if ((smem_u32(smem) & 1023u) != 0) { printf("molly attention: smem base not 1024-B aligned\n"); __trap(); }
            tma_prefetch_desc(&tma_qkv);
            tma_prefetch_desc(&tma_kv);
            mbar_init(bar_q, 1);
            for (int st = 0; st < NST; ++st) { mbar_init(&bar_kv_full[st], 1); mbar_init(&bar_kv_empty[st], 1); }
            mbar_init(bar_s_full, 1);
            mbar_init(bar_s_free, 128);
            mbar_init(bar_p_full, 128);
            mbar_init(bar_o_full, 1);
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
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + KVB;

    if (warp == 4) {
        // elect.sync, not `lane == 0`: ptxas then knows ONE thread runs this region and issues the tcgen05 / TMA instructions
        // back to back; under a lane test it wraps EACH of them in an elect / branch loop over the "active" threads (13
        // instructions and a taken branch per MMA -- the control thread, not the tensor pipe, paced the small MMAs)
        if (ATT_ELECT ? elect_one() : lane == 0) {
            // ---------------- control thread: TMA producer + MMA issuer ----------------
            constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BLOCK, KVB, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BLOCK, D, false, true);      // B (= V) is MN-major
            const uint32_t s_q = smem_u32(smem + Cfg::OFF_Q), s_k = smem_u32(smem + Cfg::OFF_K);
            const uint32_t s_v = smem_u32(smem + Cfg::OFF_V), s_p = smem_u32(smem + Cfg::OFF_P);
            auto next_item = [&](int item) {
                while (item < total && kv_info[2 * (item / (nqb * heads))] <= 0) item += gridDim.x;
                return item;
            };
            auto load_q = [&](const Item& w) {
                mbar_arrive_expect_tx(bar_q, B::TILE_BYTES);
#pragma unroll
                for (int b = 0; b < B::NBOX; ++b)
                    tma_load_2d(smem + Cfg::OFF_Q + b * B::BOX_BYTES, &tma_qkv, bar_q, w.head * D + b * B::BOX_D,
                                w.n * k_tokens + w.q0);
            };
            int l_item = next_item(blockIdx.x), l_j = 0, g_load = 0;
            Item lw = decode(l_item < total ? l_item : 0);
            auto load_next_kv = [&]() {                      // stream block g_load -> stage g_load % NST (caller: stage is free)
                if (l_item >= total) return;
                const int stg = g_load % NST, row = lw.n * k_tokens + l_j * KVB;
                mbar_arrive_expect_tx(&bar_kv_full[stg], 2 * B::KV_TILE_BYTES);
#pragma unroll
                for (int b = 0; b < B::NBOX; ++b) {
                    tma_load_2d(smem + Cfg::OFF_K + stg * B::KV_TILE_BYTES + b * B::KV_BOX_BYTES, &tma_kv, &bar_kv_full[stg],
                                h + lw.head * D + b * B::BOX_D, row);
                    tma_load_2d(smem + Cfg::OFF_V + stg * B::KV_TILE_BYTES + b * B::KV_BOX_BYTES, &tma_kv, &bar_kv_full[stg],
                                2 * h + lw.head * D + b * B::BOX_D, row);
                }
                ++g_load;
                if (++l_j == lw.nkv) {
                    l_item = next_item(l_item + gridDim.x);
                    l_j = 0;
                    if (l_item < total) lw = decode(l_item);
                }
            };
            auto issue_s = [&](int g) {
                const int st = g % NST;
                mbar_wait(&bar_kv_full[st], (g / NST) & 1);
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < D / 16; ++s) {
                    const uint32_t off = ((s * 16) / B::BOX_D) * B::BOX_BYTES + ((s * 16) % B::BOX_D) * 2;
                    const uint32_t koff = ((s * 16) / B::BOX_D) * B::KV_BOX_BYTES + ((s * 16) % B::BOX_D) * 2;
                    const uint64_t qd = make_smem_desc(s_q + off, 16, 8 * B::ROW_BYTES, B::LAYOUT);
                    const uint64_t kd = make_smem_desc(s_k + st * B::KV_TILE_BYTES + koff, 16, 8 * B::ROW_BYTES, B::LAYOUT);
                    umma_bf16_ss(tmem_s, qd, kd, idesc_s, s != 0);
                }
                umma_commit(bar_s_full);
            };
            int c_item = l_item, it = 0, g = 0;
            if (c_item < total) {
                load_q(lw);
                for (int i = 0; i < NST; ++i) load_next_kv();
                mbar_wait(bar_q, 0);
                issue_s(0);
            }
            while (c_item < total) {
                const Item w = decode(c_item);
                const int nxt = next_item(c_item + gridDim.x);
                // S(g) of this item's first block is already issued.  The softmax prologue puts it in registers:
                mbar_wait(bar_s_free, g & 1);
                tc_fence_after();
                if (w.nkv > 1) issue_s(g + 1);
                else if (nxt < total) load_q(decode(nxt));   // Q is dead once the item's last S is consumed
                for (int j = 0; j < w.nkv; ++j, ++g) {
                    const int st = g % NST;
                    // (1) S runs one block ahead of PV: s_free(g+1) arrives in the middle of the softmax step of block g
                    if (j + 1 < w.nkv) {
                        mbar_wait(bar_s_free, (g + 1) & 1);
                        tc_fence_after();
                        if (j + 2 < w.nkv) issue_s(g + 2);
                        else if (nxt < total) load_q(decode(nxt));
                    }
                    // (2) O += P(g) V(g)
                    mbar_wait(bar_p_full, g & 1);
                    tc_fence_after();
#pragma unroll
                    for (int s = 0; s < KVB / 16; ++s) {
                        const uint64_t pd = make_smem_desc(s_p + (s & 3) * 32, 16, 1024, kLayoutSW128);
                        const uint64_t vd = make_smem_desc(s_v + st * B::KV_TILE_BYTES + s * 16 * B::ROW_BYTES,
                                                           B::KV_BOX_BYTES, 8 * B::ROW_BYTES, B::LAYOUT);
                        umma_bf16_ss(tmem_o, pd, vd, idesc_pv, (j | s) != 0);
                    }
                    umma_commit(&bar_kv_empty[st]);
                    if (j == w.nkv - 1) {
                        umma_commit(bar_o_full);
                        if (nxt < total) {                   // first S of the next item, under this item's epilogue
                            mbar_wait(bar_q, (it + 1) & 1);
                            issue_s(g + 1);
                        }
                    }
                    // (3) refill the stage with stream block g + NST once PV(g) has drained it
                    if (l_item < total) {
                        mbar_wait(&bar_kv_empty[st], (g / NST) & 1);
                        load_next_kv();
                    }
                }
                ++it;
                c_item = nxt;
            }
        }
    } else {
        // ---------------- softmax warps: thread r owns query row r; two 64-key blocks live in registers ----------------
        const int r = threadIdx.x;
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        uint8_t* p_row = smem + Cfg::OFF_P + (r >> 3) * 1024 + (r & 7) * 128;
        const int sw = r & 7;
        int it = 0, g = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const Item w = decode(item);
            const bool interior = w.n_nonpad != w.kvl;
            const long long row_base = static_cast<long long>(w.n) * k_tokens;
            const bool row_ok = w.q0 + r < k_tokens;
            __nv_bfloat16* orow = out + (row_base + w.q0 + r) * h + w.head * D;
            if (w.nkv == 0) {                                 // all-pad sequence: the reference never encodes one
                if (row_ok)
                    for (int i = 0; i < D / 8; ++i) reinterpret_cast<uint4*>(orow)[i] = make_uint4(0, 0, 0, 0);
                continue;
            }
            float m_run = -CUDART_INF_F, l_run = 0.f;
            float a[KVB], b[KVB];                             // S / P of the current and of the next block

            // one pipeline step: `cur` holds S(g) (masked, max already folded into m_run / alpha); `nxt` receives S(g+1)
            auto step = [&](float (&cur)[KVB], float (&nxt)[KVB], int j, float alpha) -> float {
                const bool has_next = j + 1 < w.nkv;
                uint32_t* nraw = reinterpret_cast<uint32_t*>(nxt);
                if (has_next) {                               // S(g+1): tcgen05.ld in flight under the exp2 phase
                    mbar_wait(bar_s_full, (g + 1) & 1);
                    tc_fence_after();
                    tmem_ld32(tmem_s + lane_addr, nraw);
                    tmem_ld32(tmem_s + lane_addr + 32, nraw + 32);
                }
                const float m_use = (m_run == -CUDART_INF_F) ? 0.f : m_run;
                const uint64_t sc2 = pack_f32x2(LOG2E, LOG2E), nm2 = pack_f32x2(-m_use, -m_use);
                uint64_t sum2[2] = {pack_f32x2(0.f, 0.f), pack_f32x2(0.f, 0.f)};
#pragma unroll
                for (int i = 0; i < KVB; i += 4) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        float x0, x1;
                        unpack_f32x2(fma_f32x2(pack_f32x2(cur[i + 2 * u], cur[i + 2 * u + 1]), sc2, nm2), x0, x1);
                        if (((i >> 1) + u) % 4 < POLY) {
                            exp2_poly_pair(x0, x1);
                            cur[i + 2 * u] = x0;
                            cur[i + 2 * u + 1] = x1;
                        } else {
                            cur[i + 2 * u] = ex2(x0);
                            cur[i + 2 * u + 1] = ex2(x1);
                        }
                        sum2[u] = add_f32x2(sum2[u], pack_f32x2(cur[i + 2 * u], cur[i + 2 * u + 1]));
                    }
                }
                float sa, sb, sc, sd;
                unpack_f32x2(sum2[0], sa, sb);
                unpack_f32x2(sum2[1], sc, sd);
                l_run = l_run * alpha + ((sa + sb) + (sc + sd));
                if (has_next) {                               // S(g+1) is in registers: the control thread may issue S(g+2)
                    tmem_ld_wait_regs32(nraw);
                    tmem_ld_wait_regs32(nraw + 32);
                    tc_fence_before();
                    mbar_arrive(bar_s_free);
                }
                if (j > 0) {                                  // PV(g-1) drained the P buffer and finished O
                    mbar_wait(&bar_kv_empty[(g - 1) % NST], ((g - 1) / NST) & 1);
                    tc_fence_after();
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {                 // P -> smem, bf16, K-major, 128-B swizzle
                    uint4 u;
                    u.x = pack_bf16x2(cur[c * 8 + 0], cur[c * 8 + 1]);
                    u.y = pack_bf16x2(cur[c * 8 + 2], cur[c * 8 + 3]);
                    u.z = pack_bf16x2(cur[c * 8 + 4], cur[c * 8 + 5]);
                    u.w = pack_bf16x2(cur[c * 8 + 6], cur[c * 8 + 7]);
                    *reinterpret_cast<uint4*>(p_row + ((c ^ sw) << 4)) = u;
                }
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
                ++g;
                // off the chain: mask + row max of the next block, while the tensor core runs O += P V
                float alpha_next = 1.0f;
                if (has_next)
                    alpha_next = pipe_update_max(
                        pipe_mask_max(nxt, interior, key_mask, row_base, (j + 1) * KVB, w.kvl, k_tokens), m_run);
                return alpha_next;
            };

            // prologue: S of the item's first block
            mbar_wait(bar_s_full, g & 1);
            tc_fence_after();
            {
                uint32_t* raw = reinterpret_cast<uint32_t*>(a);
                tmem_ld32(tmem_s + lane_addr, raw);
                tmem_ld32(tmem_s + lane_addr + 32, raw + 32);
                tmem_ld_wait_regs32(raw);
                tmem_ld_wait_regs32(raw + 32);
            }
            tc_fence_before();
            mbar_arrive(bar_s_free);
            float alpha = pipe_update_max(pipe_mask_max(a, interior, key_mask, row_base, 0, w.kvl, k_tokens), m_run);
            for (int j = 0; j < w.nkv; j += 2) {
                alpha = step(a, b, j, alpha);
                if (j + 1 < w.nkv) alpha = step(b, a, j + 1, alpha);
            }
            // epilogue: O / l -> bf16 -> HBM
            mbar_wait(bar_o_full, it & 1);
            ++it;
            tc_fence_after();
            const float inv_l = 1.0f / l_run;
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
            tc_fence_before();                                // O is read: the next item's PV may overwrite it
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

long long* g_attn_debug = nullptr;     // timeline buffer of -DATT_TIMELINE builds (molly_attention_debug); nullptr in production

// Grid of the item-streaming kernel: one CTA per resident slot (MOLLY_ATTN_STREAM=0: one CTA per item, for A/B timing).
int attention_grid(int total_items, int ctas_per_sm) {
    static int stream_items = -1;
    if (stream_items < 0) {
        const char* e = getenv("MOLLY_ATTN_STREAM");
        stream_items = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const int slots = device_sm_count() * ctas_per_sm;
    return (stream_items && total_items > slots) ? slots : total_items;
}

template <int D, int KVB>
int launch_attention_kvb(const AttnMaps& maps, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                         const uint8_t* key_mask, void* out, float* lse, cudaStream_t stream) {
    using Cfg = AttnCfg<D, KVB>;
    auto kernel = attention_kernel<D, KVB, 0>;
    static bool configured = false;
    if (!configured) {
        MOLLY_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int grid = attention_grid(n_seq * heads * ((k_tokens + ATT_BLOCK - 1) / ATT_BLOCK), Cfg::MIN_CTAS);
    {
        prof_attention_work(kv_info, n_seq, k_tokens, h, 4.0, stream);      // work = 4 h K sum(kv_len), known on the device only
        ProfScope prof(PF_ATTENTION, 4.0 * n_seq * k_tokens * static_cast<double>(k_tokens) * h, stream);
        kernel<<<grid, ATT_THREADS, Cfg::SMEM_BYTES, stream>>>(maps.q, maps.kv64, n_seq, heads, k_tokens, h, kv_info,
                                                               key_mask, static_cast<__nv_bfloat16*>(out), lse);
    }
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

// keys per KV block: 64 puts three CTAs on an SM at head_dim <= 64 (MOLLY_ATTN_KVB = 64 | 128 overrides)
int attention_kvb(int d) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("MOLLY_ATTN_KVB");
        forced = e == nullptr ? 0 : atoi(e);
    }
    if (forced == 64 && d <= 64) return 64;
    if (forced == 128) return 128;
    return (ATTN_KVB64_DEFAULT && d <= 64) ? 64 : 128;
}

template <int D>
int launch_attention_pipe(const AttnMaps& maps, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                          const uint8_t* key_mask, void* out, cudaStream_t stream) {
    using Cfg = AttnPipeCfg<D>;
    static int poly = -1;
    if (poly < 0) {
        const char* e = getenv("MOLLY_ATTN_POLY");
        poly = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : ATTN_POLY_DEFAULT;
    }
    auto kernel = poly == 0 ? attention_pipe_kernel<D, 0>
                            : (poly == 1 ? attention_pipe_kernel<D, 1> : attention_pipe_kernel<D, 2>);
    static bool configured = false;
    if (!configured) {
        MOLLY_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int grid = attention_grid(n_seq * heads * ((k_tokens + ATT_BLOCK - 1) / ATT_BLOCK), 2);
    {
        prof_attention_work(kv_info, n_seq, k_tokens, h, 4.0, stream);      // work = 4 h K sum(kv_len), known on the device only
        ProfScope prof(PF_ATTENTION, 4.0 * n_seq * k_tokens * static_cast<double>(k_tokens) * h, stream);
        kernel<<<grid, ATT_THREADS, Cfg::SMEM_BYTES, stream>>>(maps.q, maps.kv64, n_seq, heads, k_tokens, h, kv_info,
                                                               key_mask, static_cast<__nv_bfloat16*>(out));
    }
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

bool attention_pipe_enabled() {         // MOLLY_ATTN_PIPE = 0 | 1 overrides the default
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MOLLY_ATTN_PIPE");
        v = e == nullptr ? (ATTN_PIPE_DEFAULT ? 1 : 0) : (e[0] != '0');
    }
    return v == 1;
}

template <int D>
int launch_attention(const AttnMaps& maps, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                     const uint8_t* key_mask, void* out, float* lse, cudaStream_t stream) {
    const CUtensorMap& tm = maps.q;
    if constexpr (D <= 64) {
        if (attention_pipe_enabled() && lse == nullptr)
            return launch_attention_pipe<D>(maps, n_seq, k_tokens, h, heads, kv_info, key_mask, out, stream);
        if (attention_kvb(D) == 64 && lse == nullptr)
            return launch_attention_kvb<D, 64>(maps, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
    }
    using Cfg = AttnCfg<D>;
    static int poly = -1;             // pairs out of 4 whose exp2 runs on the FMA pipe (MOLLY_ATTN_POLY = 0 | 1 | 2)
    if (poly < 0) {
        const char* e = getenv("MOLLY_ATTN_POLY");
        poly = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : ATTN_POLY_DEFAULT;
    }
    auto kernel = lse != nullptr ? attention_kernel<D, 128, 0, true>
                                 : (poly == 0 ? attention_kernel<D, 128, 0>
                                              : (poly == 1 ? attention_kernel<D, 128, 1> : attention_kernel<D, 128, 2>));
    static bool configured = false;
    if (!configured) {
        MOLLY_CUDA(cudaFuncSetAttribute(attention_kernel<D, 128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention_kernel<D, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention_kernel<D, 128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        MOLLY_CUDA(cudaFuncSetAttribute(attention_kernel<D, 128, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
        configured = true;
    }
    const int grid = attention_grid(n_seq * heads * ((k_tokens + ATT_BLOCK - 1) / ATT_BLOCK), Cfg::MIN_CTAS);
    {   // dense-equivalent work 4*n*K*K*h (exact when every sequence is full length)
        prof_attention_work(kv_info, n_seq, k_tokens, h, 4.0, stream);      // work = 4 h K sum(kv_len), known on the device only
        ProfScope prof(PF_ATTENTION, 4.0 * n_seq * k_tokens * static_cast<double>(k_tokens) * h, stream);
        kernel<<<grid, ATT_THREADS, Cfg::SMEM_BYTES, stream>>>(tm, tm, n_seq, heads, k_tokens, h, kv_info, key_mask,
                                                               static_cast<__nv_bfloat16*>(out), lse);
    }
    count_launch();
    MOLLY_CUDA(cudaGetLastError());
    return MOLLY_OK;
}

}  // namespace

void attention2_set_debug(long long* buf);
void attention_set_debug(long long* buf) {
    g_attn_debug = buf;
    attention2_set_debug(buf);
#ifdef ATT_TIMELINE
    cudaMemcpyToSymbol(d_attn_tl, &buf, sizeof(buf));
#endif
}

int attention_make_map(AttnMaps* maps, const void* qkv, int rows, int h, int heads) {
    const int d = h / heads;
    MOLLY_CHECK(d == 16 || d == 32 || d == 64 || d == 128, MOLLY_ERR_UNSUPPORTED,
                "attention: head_dim %d not in {16,32,64,128}", d);
    const int box_d = d < 64 ? d : 64;
    int rc = make_tma_2d(&maps->q, qkv, rows, 3 * h, 3 * h, ATT_BLOCK, box_d, 2);
    if (rc) return rc;
    return make_tma_2d(&maps->kv64, qkv, rows, 3 * h, 3 * h, 64, box_d, 2);
}

int attention_launch(const AttnMaps& tqkv, int n_seq, int k_tokens, int h, int heads, const int32_t* kv_info,
                     const uint8_t* key_mask, void* out, cudaStream_t stream, float* lse) {
    MOLLY_CHECK(n_seq > 0 && k_tokens > 0 && heads > 0 && h % heads == 0, MOLLY_ERR_INVALID,
                "attention: bad shape n_seq=%d k=%d h=%d heads=%d", n_seq, k_tokens, h, heads);
    MOLLY_CHECK(static_cast<long long>(n_seq) * k_tokens < (1ll << 31), MOLLY_ERR_UNSUPPORTED,
                "attention: n_seq*k_tokens exceeds int32 TMA coordinates");
    MOLLY_CHECK(n_seq <= 65535 && heads <= 65535, MOLLY_ERR_UNSUPPORTED, "attention: grid too large");
    if (attention2_enabled(h / heads))
        return attention2_launch(tqkv, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
    switch (h / heads) {
        case 16: return launch_attention<16>(tqkv, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        case 32: return launch_attention<32>(tqkv, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        case 64: return launch_attention<64>(tqkv, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        case 128: return launch_attention<128>(tqkv, n_seq, k_tokens, h, heads, kv_info, key_mask, out, lse, stream);
        default: MOLLY_CHECK(false, MOLLY_ERR_UNSUPPORTED, "attention: head_dim %d unsupported", h / heads);
    }
}

}  // namespace molly
