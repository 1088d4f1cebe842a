// affine_packed16.cuh -- score pass of the affine:local lattice fill with TWO
// lattices per warp, one in each 16-bit half of every register (DPX S16x2).
//
// Same recurrence, same closed-model order and the same END rule as
// affine_fill_kernel<R, false, END_ANYWHERE, SCORE_PRMT> (affine_systolic.cuh;
// reference: generated optimal:affine:local find score / find region,
// src/c4/viterbi.c:1638-1727, semantics viterbi.c:655-837), and bit-identical
// results: the int32 max-plus values of a LOCAL lattice are bounded,
//     0 <= M <= max_sub * min(Q, T),    D, I >= 2 * gap_open + gap_extend,
// so when max_sub * (min(Q,T) + 1) fits 15 bits no 16-bit add can wrap and the
// halfword arithmetic IS the int32 arithmetic.  The host (affine_create) sends a
// lattice here only when that bound, Q + 1 <= 32 R (one sweep) and <= 4 query
// symbol classes hold; everything else takes the int32 kernel.
//
// Mapping: one CTA = one warp = lattices 2b (low halves) and 2b+1 (high halves).
// VIADDMNMX.S16x2 / VIMNMX.S16x2.RELU / VIMNMX3.S16x2 each advance both lattices,
// and ONE prmt builds both sign-extended substitution scores: the 8-byte pool is
// {column of A's target symbol (4 classes), column of B's target symbol}, the
// per-row selector picks (class_A, sign, 4 + class_B, sign).
//
// Padding is by input conditioning: rows below a query select "sign of a pool
// byte" for both bytes (score' = 0 or -1, i.e. s = gap_open or gap_open - 1 < 0)
// and columns right of a target use the "no symbol" column (score' = 0).  Every
// move into or inside the padding is strictly negative, so a padded cell with
// value v > 0 descends from a real cell with value > v that precedes it in the
// reference's scan order (target outer, query inner): padding can never be the
// first maximum (viterbi.c:778-791), and real cells never read padded ones.
#pragma once
#include "affine_systolic.cuh"

namespace c4b {

constexpr uint32_t kNeg16x2 = 0xC000C000u;  // -16384 | -16384: "not reachable"
constexpr uint32_t kMin16x2 = 0x80008000u;  // -32768 | -32768

__device__ __forceinline__ uint32_t pack16(int v) {
    return ((uint32_t)v & 0xFFFFu) * 0x10001u;
}
// 0xFFFF in every halfword whose sign bit is set, else 0 (one PRMT)
__device__ __forceinline__ uint32_t sign_mask16x2(uint32_t v) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(0u), "r"(0xBB99u));
    return d;
}
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(short)(v >> 16); }

template <int R>
__global__ void __launch_bounds__(32, 12)
affine_fill16_kernel(const AffPair *__restrict__ pairs, AffOut *__restrict__ outs, const int n,
                     const AffModel mdl, const void *__restrict__ score_table) {
    __shared__ uint32_t xt4[25];
    const int lane = threadIdx.x;
    const int ia = 2 * blockIdx.x, ib = min(ia + 1, n - 1);
    const AffPair PA = pairs[ia], PB = pairs[ib];
    const int QA = PA.Q, TA = PA.T, QB = PB.Q, TB_ = PB.T;
    const int T = max(TA, TB_);
    if (lane < 25) xt4[lane] = reinterpret_cast<const uint2 *>(score_table)[lane].x;  // classes 0..3
    __syncwarp();

    const int open = mdl.openD;
    const uint32_t open2 = pack16(open), extD2 = pack16(mdl.extD), extI2 = pack16(mdl.extI);
    const int nsteps = T + 1 + 31;
    const int row0 = lane * R;

    uint32_t sel[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = row0 + r;  // lattice row; consumes query symbol i-1
        uint32_t sa = 0x88u, sb = 0xCCu;  // padding: sign of pool byte 0 / byte 4
        if (i >= 1 && i <= QA) { const uint32_t c = PA.q[i - 1]; sa = c | ((c | 8u) << 4); }
        if (i >= 1 && i <= QB) { const uint32_t c = 4u + PB.q[i - 1]; sb = c | ((c | 8u) << 4); }
        sel[r] = sa | (sb << 8);
    }
    uint32_t Mp[R], Dp[R];  // G = M + open and D of the previous column, both lattices
#pragma unroll
    for (int r = 0; r < R; ++r) {
        Mp[r] = kNeg16x2;
        Dp[r] = kNeg16x2;
    }
    uint32_t topM = kNeg16x2, topI = kNeg16x2, topMprev = kNeg16x2;
    uint32_t in_code = kTargetNone | (kTargetNone << 8), code0 = in_code;

    // per lattice: first strict maximum of END in (target outer, query inner) order
    uint32_t best2 = kMin16x2;
    int bjA = 0, biA = 0, bjB = 0, biB = 0;
    uint32_t pend = kMin16x2;
    int pend_j = 0;
    auto settle_pending = [&]() {
        const uint32_t nb = __vimax3_s16x2(best2, pend, pend);
        if (nb != best2) {  // some half improved strictly (columns arrive in increasing j)
            if (lo16(pend) > lo16(best2)) {
                int bi = 0;
                bool found = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!found && ((Mp[r] ^ pend) & 0xFFFFu) == 0) { bi = row0 + r; found = true; }
                biA = bi;
                bjA = pend_j;
            }
            if (hi16(pend) > hi16(best2)) {
                int bi = 0;
                bool found = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!found && ((Mp[r] ^ pend) >> 16) == 0) { bi = row0 + r; found = true; }
                biB = bi;
                bjB = pend_j;
            }
            best2 = nb;
        }
        pend = kMin16x2;
    };

    auto step = [&](const int s, auto ALL) {
        constexpr bool all_active = decltype(ALL)::value;
        const int j = s - lane;
        settle_pending();
        const uint32_t code = (lane == 0) ? code0 : in_code;
        {   // column s+1 of both lattices (uniform addresses, lane 0 consumes them)
            const uint32_t ca = (s + 1 <= TA) ? (uint32_t)PA.t[s] : (uint32_t)kTargetNone;
            const uint32_t cb = (s + 1 <= TB_) ? (uint32_t)PB.t[s] : (uint32_t)kTargetNone;
            code0 = ca | (cb << 8);
        }
        uint32_t botM = kNeg16x2, botI = kNeg16x2;
        if (all_active || (j >= 0 && j <= T)) {
            const uint32_t Xa = xt4[code & 0xFFu], Xb = xt4[code >> 8];
            uint32_t cm = kMin16x2;
            // phase A, bottom-up, rows independent: D of this column, max(match, D)
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                uint32_t sc;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sc) : "r"(Xa), "r"(Xb), "r"(sel[r]));
                const uint32_t diag = (r == 0) ? topMprev : Mp[r - 1];
                Dp[r] = __viaddmax_s16x2(Dp[r], extD2, Mp[r]);
                Mp[r] = __viaddmax_s16x2(diag, sc, Dp[r]);
            }
            // phase B, top-down: the vertical chain I -> M -> G
            uint32_t upM = topM, upI = topI;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t Iv = __viaddmax_s16x2(upI, extI2, upM);
                const uint32_t Mv = __vimax_s16x2_relu(Mp[r], Iv);       // START's 0 is the RELU
                const uint32_t Gv = __viaddmax_s16x2(Mv, open2, kMin16x2);  // M + open per half
                Mp[r] = Gv;
                upM = Gv;
                upI = Iv;
                if (r & 1) cm = __vimax3_s16x2(cm, Gv, Mp[r - 1]);
            }
            botM = upM;
            botI = upI;
            topMprev = topM;
            pend = cm;
            pend_j = j;
        }
        const uint32_t nM = __shfl_up_sync(0xffffffffu, botM, 1);
        const uint32_t nI = __shfl_up_sync(0xffffffffu, botI, 1);
        const uint32_t nC = __shfl_up_sync(0xffffffffu, code, 1);
        if (lane > 0) {
            topM = nM;
            topI = nI;
            in_code = nC;
        }
    };

    const int fill_end = min(31, nsteps);
    const int steady_end = max(fill_end, min(T + 1, nsteps));
    int s = 0;
    for (; s < fill_end; ++s) step(s, std::false_type{});
    for (; s < steady_end; ++s) step(s, std::true_type{});
    for (; s < nsteps; ++s) step(s, std::false_type{});
    settle_pending();
    __syncwarp();

    // lexicographic warp reduction per lattice: max score, then min j, then min i
    int bA = lo16(best2), bB = hi16(best2);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bA, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bjA, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biA, off);
            if ((ob > bA) || (ob == bA && (oj < bjA || (oj == bjA && oi < biA)))) { bA = ob; bjA = oj; biA = oi; }
        }
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bB, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bjB, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biB, off);
            if ((ob > bB) || (ob == bB && (oj < bjB || (oj == bjB && oi < biB)))) { bB = ob; bjB = oj; biB = oi; }
        }
    }
    if (lane == 0) {
        AffOut o;
        o.best = bA - open;  // tracked as G = M + open
        o.end_i = biA;
        o.end_j = bjA;
        o.flags = 0;
        outs[PA.out_index] = o;
        if (ia + 1 < n) {
            o.best = bB - open;
            o.end_i = biB;
            o.end_j = bjB;
            outs[PB.out_index] = o;
        }
    }
}


// -----------------------------------------------------------------------------
// affine_fill16u_kernel: the same two-lattices-per-warp score pass, re-balanced
// over the SM's pipes (measured with tools/ubench/dpx_rates.cu on B200: the
// three-input DPX forms VIADDMNMX / VIMNMX3 and PRMT issue every ~2.4 cycles per
// scheduler on the ALU pipe, two-input VIMNMX every ~1.1, IMAD every ~2.2 on the
// FMA pipe).  Two changes against affine_fill16_kernel:
//
//  1. OFFSET-BINARY halves: every halfword holds value + kBias16 and is compared
//     UNSIGNED.  All stored values are >= kBias16 - 16000 > |any penalty| and the
//     score' entries are >= 0 (host-checked: s >= gap_open), so "+ penalty" and
//     "+ score'" on both halves is ONE 32-bit integer add with no carry or borrow
//     across the halves -- an IMAD by a run-time 1 on the FMA pipe -- and the max
//     is the cheap two-input VIMNMX.U16x2.  START's 0 is max(., kBias16).
//  2. SHORT VERTICAL CHAIN: with G~ = max(match, D, START) + open (independent of the
//     row above),  I(r+1) = max(I(r) + ext, M(r) + open)
//                         = max(I(r) + ext, G~(r), I(r) + open) = max(I(r) + ext, G~(r))
//     because open <= ext (host-checked).  The only loop-carried dependency down the
//     rows is ONE VIADDMNMX.U16x2 per row; G = max(G~, I + open) is off the chain.
// Per packed row: ALU 10.5 cycles (PRMT, VIMNMX, VIMNMX3, VIADDMNMX, VIMNMX, half a
// VIMNMX3), FMA 8.6 (4 IMAD), against 14.1 ALU cycles of the signed kernel.
constexpr uint32_t kBias16 = 0x4000u;
constexpr uint32_t kBias16x2 = 0x40004000u;   // START / RELU floor: true 0
constexpr uint32_t kNegU16x2 = 0x01800180u;   // true -16000: "not reachable"

__device__ __forceinline__ uint32_t imad_add(uint32_t a, int one, uint32_t k) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"((uint32_t)one), "r"(k));
    return d;
}

// one sweep, one warp per pair of lattices (the metric configuration): 168 registers, 12 CTAs per SM
template <int R>
__global__ void __launch_bounds__(32, 12)
affine_fill16u_kernel(const AffPair *__restrict__ pairs, AffOut *__restrict__ outs, const int n,
                      const AffModel mdl, const void *__restrict__ score_table) {
    __shared__ uint32_t xt4[25];
    const int lane = threadIdx.x;
    const int ia = 2 * blockIdx.x, ib = min(ia + 1, n - 1);
    const AffPair PA = pairs[ia], PB = pairs[ib];
    const int QA = PA.Q, TA = PA.T, QB = PB.Q, TB_ = PB.T;
    const int T = max(TA, TB_);
    if (lane < 25) xt4[lane] = reinterpret_cast<const uint2 *>(score_table)[lane].x;  // classes 0..3, all >= 0
    __syncwarp();

    const int open = mdl.openD, one = mdl.one;
    // x + {p, p} for a penalty p < 0 and halves >= |p|: one 32-bit add of p * 0x10001
    const uint32_t openK = (uint32_t)(open * 0x10001), extDK = (uint32_t)(mdl.extD * 0x10001);
    const uint32_t extI2 = pack16(mdl.extI);  // per-half operand of VIADDMNMX.U16x2 (wraps per half)
    const int nsteps = T + 1 + 31;
    const int row0 = lane * R;

    uint32_t sel[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = row0 + r;
        uint32_t sa = 0x88u, sb = 0xCCu;  // padding: sign of a (non-negative) pool byte = 0
        if (i >= 1 && i <= QA) { const uint32_t c = PA.q[i - 1]; sa = c | ((c | 8u) << 4); }
        if (i >= 1 && i <= QB) { const uint32_t c = 4u + PB.q[i - 1]; sb = c | ((c | 8u) << 4); }
        sel[r] = sa | (sb << 8);
    }
    uint32_t Mp[R], Dp[R];  // G = M + open and D of the previous column (offset binary)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        Mp[r] = kNegU16x2;
        Dp[r] = kNegU16x2;
    }
    uint32_t topM = kNegU16x2, topI = kNegU16x2, topMprev = kNegU16x2;
    uint32_t in_code = kTargetNone | (kTargetNone << 8), code0 = in_code;

    uint32_t best2 = 0u;  // below every stored value
    int bjA = 0, biA = 0, bjB = 0, biB = 0;
    uint32_t pend = 0u;
    int pend_j = 0;
    auto settle_pending = [&]() {
        const uint32_t nb = __vmaxu2(best2, pend);
        if (nb != best2) {  // some half improved strictly (columns arrive in increasing j)
            if ((pend & 0xFFFFu) > (best2 & 0xFFFFu)) {
                int bi = 0;
                bool found = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!found && ((Mp[r] ^ pend) & 0xFFFFu) == 0) { bi = row0 + r; found = true; }
                biA = bi;
                bjA = pend_j;
            }
            if ((pend >> 16) > (best2 >> 16)) {
                int bi = 0;
                bool found = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!found && ((Mp[r] ^ pend) >> 16) == 0) { bi = row0 + r; found = true; }
                biB = bi;
                bjB = pend_j;
            }
            best2 = nb;
        }
        pend = 0u;
    };

    auto step = [&](const int s, auto ALL) {
        constexpr bool all_active = decltype(ALL)::value;
        const int j = s - lane;
        settle_pending();
        const uint32_t code = (lane == 0) ? code0 : in_code;
        {
            const uint32_t ca = (s + 1 <= TA) ? (uint32_t)PA.t[s] : (uint32_t)kTargetNone;
            const uint32_t cb = (s + 1 <= TB_) ? (uint32_t)PB.t[s] : (uint32_t)kTargetNone;
            code0 = ca | (cb << 8);
        }
        uint32_t botM = kNegU16x2, botI = kNegU16x2;
        if (all_active || (j >= 0 && j <= T)) {
            const uint32_t Xa = xt4[code & 0xFFu], Xb = xt4[code >> 8];
            uint32_t cm = 0u;
            // phase A, bottom-up, rows independent: D, then G~ = max(match, D, START) + open
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                uint32_t sc;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sc) : "r"(Xa), "r"(Xb), "r"(sel[r]));
                const uint32_t diag = (r == 0) ? topMprev : Mp[r - 1];
                Dp[r] = __vmaxu2(imad_add(Dp[r], one, extDK), Mp[r]);
                const uint32_t x = __vimax3_u16x2(imad_add(diag, one, sc), Dp[r], kBias16x2);
                Mp[r] = imad_add(x, one, openK);
            }
            // phase B, top-down: I chain (one op per row), G = max(G~, I + open) off the chain
            uint32_t Iv = __viaddmax_u16x2(topI, extI2, topM);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t gt = Mp[r];
                const uint32_t Gv = __vmaxu2(gt, imad_add(Iv, one, openK));
                Mp[r] = Gv;
                if (r & 1) cm = __vimax3_u16x2(cm, Gv, Mp[r - 1]);
                if (r + 1 < R) Iv = __viaddmax_u16x2(Iv, extI2, gt);
            }
            botM = Mp[R - 1];
            botI = Iv;
            topMprev = topM;
            pend = cm;
            pend_j = j;
        }
        const uint32_t nM = __shfl_up_sync(0xffffffffu, botM, 1);
        const uint32_t nI = __shfl_up_sync(0xffffffffu, botI, 1);
        const uint32_t nC = __shfl_up_sync(0xffffffffu, code, 1);
        if (lane > 0) {
            topM = nM;
            topI = nI;
            in_code = nC;
        }
    };

    const int fill_end = min(31, nsteps);
    const int steady_end = max(fill_end, min(T + 1, nsteps));
    int s = 0;
    for (; s < fill_end; ++s) step(s, std::false_type{});
    for (; s < steady_end; ++s) step(s, std::true_type{});
    for (; s < nsteps; ++s) step(s, std::false_type{});
    settle_pending();
    __syncwarp();

    int bA = (int)(best2 & 0xFFFFu), bB = (int)(best2 >> 16);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bA, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bjA, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biA, off);
            if ((ob > bA) || (ob == bA && (oj < bjA || (oj == bjA && oi < biA)))) { bA = ob; bjA = oj; biA = oi; }
        }
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bB, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bjB, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biB, off);
            if ((ob > bB) || (ob == bB && (oj < bjB || (oj == bjB && oi < biB)))) { bB = ob; bjB = oj; biB = oi; }
        }
    }
    if (lane == 0) {
        AffOut o;
        o.best = bA - (int)kBias16 - open;  // tracked as G = M + open, offset binary
        o.end_i = biA;
        o.end_j = bjA;
        o.flags = 0;
        outs[PA.out_index] = o;
        if (ia + 1 < n) {
            o.best = bB - (int)kBias16 - open;
            o.end_i = biB;
            o.end_j = bjB;
            outs[PB.out_index] = o;
        }
    }
}


// The same kernel for queries of several sweeps (kept separate: folding the sweep loop away at
// compile time still cost the one-sweep kernel 10 % on the B200).  MULTI is always true here.
template <int R, bool MULTI>
__device__ __forceinline__ void affine_fill16u_body(const AffPair *__restrict__ pairs, AffOut *__restrict__ outs,
                                                    const int n, const AffModel &mdl,
                                                    const void *__restrict__ score_table) {
    // blockDim.x = 32 W.  Queries longer than one sweep (32 R rows) are swept in strips,
    // and the W warps of the CTA run the strips of this PAIR of lattices concurrently as a
    // pipeline, exactly as affine_fill_kernel does (hand-off row {G, I} of both lattices in
    // lattice A's top0/top1 buffers, published through a monotone counter in shared memory).
    __shared__ uint32_t xt4[25];
    __shared__ volatile long long vprog[kAffMaxWarps];
    __shared__ int red[kAffMaxWarps][6];
    const int lane = threadIdx.x & 31, warp = MULTI ? (int)(threadIdx.x >> 5) : 0, W = MULTI ? (int)(blockDim.x >> 5) : 1;
    const int ia = 2 * blockIdx.x, ib = min(ia + 1, n - 1);
    const AffPair PA = pairs[ia], PB = pairs[ib];
    const int QA = PA.Q, TA = PA.T, QB = PB.Q, TB_ = PB.T;
    const int T = max(TA, TB_);
    if (threadIdx.x < 25) xt4[threadIdx.x] = reinterpret_cast<const uint2 *>(score_table)[threadIdx.x].x;  // classes 0..3, all >= 0
    if (MULTI && threadIdx.x < kAffMaxWarps) vprog[threadIdx.x] = 0;
    __syncthreads();

    const int open = mdl.openD, one = mdl.one;
    // x + {p, p} for a penalty p < 0 and halves >= |p|: one 32-bit add of p * 0x10001
    const uint32_t openK = (uint32_t)(open * 0x10001), extDK = (uint32_t)(mdl.extD * 0x10001);
    const uint32_t extI2 = pack16(mdl.extI);  // per-half operand of VIADDMNMX.U16x2 (wraps per half)
    const int nsteps = T + 1 + 31;
    const int pub_mask = (T >= 4096) ? 15 : 3;   // hand-off rows are published every 16 (4) columns
    const int rows_per_sweep = 32 * R;
    const int nsweeps = MULTI ? (max(QA, QB) + 1 + rows_per_sweep - 1) / rows_per_sweep : 1;

    uint32_t best2 = 0u;  // below every stored value
    int bjA = 0, biA = 0, bjB = 0, biB = 0;

    for (int sweep = warp; sweep < nsweeps; sweep += W) {
        const int row0 = sweep * rows_per_sweep + lane * R;
        const bool later_sweep = MULTI && (sweep > 0);
        const bool first_row_lane = (sweep == 0 && lane == 0);
        // within one sweep columns arrive in increasing j, so only a strictly greater score can
        // be an earlier END; a LATER sweep of this warp can tie an earlier one at a smaller j
        const bool tie_possible = MULTI && (sweep != warp);
        uint32_t sel[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = row0 + r;
            uint32_t sa = 0x88u, sb = 0xCCu;  // padding: sign of a (non-negative) pool byte = 0
            if (i >= 1 && i <= QA) { const uint32_t c = PA.q[i - 1]; sa = c | ((c | 8u) << 4); }
            if (i >= 1 && i <= QB) { const uint32_t c = 4u + PB.q[i - 1]; sb = c | ((c | 8u) << 4); }
            sel[r] = sa | (sb << 8);
        }
        uint32_t Mp[R], Dp[R];  // G = M + open and D of the previous column (offset binary)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            Mp[r] = kNegU16x2;
            Dp[r] = kNegU16x2;
        }
        const uint2 *top_in = reinterpret_cast<const uint2 *>((sweep & 1) ? PA.top0 : PA.top1);  // written by sweep-1
        uint2 *top_out = reinterpret_cast<uint2 *>((sweep & 1) ? PA.top1 : PA.top0);
        const bool write_top = MULTI && (sweep + 1 < nsweeps) && (lane == 31);
        const bool piped = MULTI && later_sweep && W > 1;
        const int wp = (sweep - 1) % W;
        const long long in_base = (long long)(sweep - 1) * (T + 1);
        long long avail = 0;
        auto wait_column = [&](int col) {
            const long long need = in_base + col + 1;
            if (avail < need) {
                while ((avail = vprog[wp]) < need) __nanosleep(40);
                __threadfence_block();
            }
        };
        uint32_t topM = kNegU16x2, topI = kNegU16x2, topMprev = kNegU16x2;
        uint32_t in_code = kTargetNone | (kTargetNone << 8), code0 = in_code;
        uint2 top0v = make_uint2(kNegU16x2, kNegU16x2);
        if (later_sweep) {
            if (piped) wait_column(0);
            top0v = __ldcg(top_in);
        }

        uint32_t pend = 0u;
        int pend_j = 0;
        auto settle_pending = [&]() {
            const uint32_t nb = __vmaxu2(best2, pend);
            // strictly greater in some half, or (later sweeps of this warp) a tie at a smaller j
            bool trig = (nb != best2);
            if (tie_possible)
                trig = trig || (pend != 0u && ((((pend ^ best2) & 0xFFFFu) == 0 && pend_j < bjA) ||
                                               (((pend ^ best2) >> 16) == 0 && pend_j < bjB)));
            if (trig) {
                const uint32_t pl = pend & 0xFFFFu, bl = best2 & 0xFFFFu, ph = pend >> 16, bh = best2 >> 16;
                if (pl > bl || (pl == bl && pl != 0u && pend_j < bjA)) {
                    int bi = 0;
                    bool found = false;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (!found && ((Mp[r] ^ pend) & 0xFFFFu) == 0) { bi = row0 + r; found = true; }
                    biA = bi;
                    bjA = pend_j;
                }
                if (ph > bh || (ph == bh && ph != 0u && pend_j < bjB)) {
                    int bi = 0;
                    bool found = false;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (!found && ((Mp[r] ^ pend) >> 16) == 0) { bi = row0 + r; found = true; }
                    biB = bi;
                    bjB = pend_j;
                }
                best2 = nb;
            }
            pend = 0u;
        };

        auto step = [&](const int s, auto ALL) {
            constexpr bool all_active = decltype(ALL)::value;
            const int j = s - lane;
            settle_pending();
            const uint32_t code = (lane == 0) ? code0 : in_code;
            if (later_sweep && lane == 0) {
                topM = top0v.x;
                topI = top0v.y;
            }
            {
                const uint32_t ca = (s + 1 <= TA) ? (uint32_t)PA.t[s] : (uint32_t)kTargetNone;
                const uint32_t cb = (s + 1 <= TB_) ? (uint32_t)PB.t[s] : (uint32_t)kTargetNone;
                code0 = ca | (cb << 8);
                if (later_sweep && s + 1 <= T) {
                    if (piped) wait_column(s + 1);
                    top0v = __ldcg(top_in + s + 1);
                }
            }
            uint32_t botM = kNegU16x2, botI = kNegU16x2;
            if (all_active || (j >= 0 && j <= T)) {
                const uint32_t Xa = xt4[code & 0xFFu], Xb = xt4[code >> 8];
                uint32_t cm = 0u;
                // phase A, bottom-up, rows independent: D, then G~ = max(match, D, START) + open
#pragma unroll
                for (int r = R - 1; r >= 0; --r) {
                    uint32_t sc;
                    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sc) : "r"(Xa), "r"(Xb), "r"(sel[r]));
                    const uint32_t diag = (r == 0) ? topMprev : Mp[r - 1];
                    Dp[r] = __vmaxu2(imad_add(Dp[r], one, extDK), Mp[r]);
                    const uint32_t x = __vimax3_u16x2(imad_add(diag, one, sc), Dp[r], kBias16x2);
                    Mp[r] = imad_add(x, one, openK);
                }
                // phase B, top-down: I chain (one op per row), G = max(G~, I + open) off the chain
                uint32_t Iv = __viaddmax_u16x2(topI, extI2, topM);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t gt = Mp[r];
                    const uint32_t Gv = __vmaxu2(gt, imad_add(Iv, one, openK));
                    Mp[r] = Gv;
                    if (r & 1) cm = __vimax3_u16x2(cm, Gv, Mp[r - 1]);
                    if (r + 1 < R) Iv = __viaddmax_u16x2(Iv, extI2, gt);
                }
                botM = Mp[R - 1];
                botI = Iv;
                topMprev = topM;
                pend = cm;
                pend_j = j;
                if (write_top) {
                    top_out[j] = make_uint2(botM, botI);
                    if (MULTI && W > 1 && ((j & pub_mask) == pub_mask || j == T)) {   // (groups: see affine_fill_kernel)
                        __threadfence_block();
                        vprog[warp] = (long long)sweep * (T + 1) + j + 1;
                    }
                }
            }
            if (first_row_lane) { topM = kNegU16x2; topI = kNegU16x2; }
            const uint32_t nM = __shfl_up_sync(0xffffffffu, botM, 1);
            const uint32_t nI = __shfl_up_sync(0xffffffffu, botI, 1);
            const uint32_t nC = __shfl_up_sync(0xffffffffu, code, 1);
            if (lane > 0) {
                topM = nM;
                topI = nI;
                in_code = nC;
            }
        };

        const int fill_end = min(31, nsteps);
        const int steady_end = max(fill_end, min(T + 1, nsteps));
        int s = 0;
        for (; s < fill_end; ++s) step(s, std::false_type{});
        for (; s < steady_end; ++s) step(s, std::true_type{});
        for (; s < nsteps; ++s) step(s, std::false_type{});
        settle_pending();
        __syncwarp();
    }

    int bA = (int)(best2 & 0xFFFFu), bB = (int)(best2 >> 16);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bA, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bjA, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biA, off);
            if ((ob > bA) || (ob == bA && (oj < bjA || (oj == bjA && oi < biA)))) { bA = ob; bjA = oj; biA = oi; }
        }
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bB, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bjB, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biB, off);
            if ((ob > bB) || (ob == bB && (oj < bjB || (oj == bjB && oi < biB)))) { bB = ob; bjB = oj; biB = oi; }
        }
    }
    if (MULTI && W > 1) {   // combine the warps' sweeps (idle warps carry 0 = below every stored value)
        if (lane == 0) {
            red[warp][0] = bA; red[warp][1] = bjA; red[warp][2] = biA;
            red[warp][3] = bB; red[warp][4] = bjB; red[warp][5] = biB;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 1; w < W; ++w) {
                int ob = red[w][0], oj = red[w][1], oi = red[w][2];
                if ((ob > bA) || (ob == bA && (oj < bjA || (oj == bjA && oi < biA)))) { bA = ob; bjA = oj; biA = oi; }
                ob = red[w][3]; oj = red[w][4]; oi = red[w][5];
                if ((ob > bB) || (ob == bB && (oj < bjB || (oj == bjB && oi < biB)))) { bB = ob; bjB = oj; biB = oi; }
            }
    }
    if (threadIdx.x == 0) {
        AffOut o;
        o.best = bA - (int)kBias16 - open;  // tracked as G = M + open, offset binary
        o.end_i = biA;
        o.end_j = bjA;
        o.flags = 0;
        outs[PA.out_index] = o;
        if (ia + 1 < n) {
            o.best = bB - (int)kBias16 - open;
            o.end_i = biB;
            o.end_j = bjB;
            outs[PB.out_index] = o;
        }
    }
}

// long queries (R = 32), or small batches that take fewer rows per lane to occupy the GPU
// (R = 16 / 8, host: affine_create): up to 8 warps per pair of lattices, sweeps pipelined
template <int R>
__global__ void __maxnreg__(168)   // R = 32: 3 CTAs of 4 warps, or 1 of 8, per SM
affine_fill16u_multi_kernel(const AffPair *__restrict__ pairs, AffOut *__restrict__ outs, const int n,
                            const AffModel mdl, const void *__restrict__ score_table) {
    affine_fill16u_body<R, true>(pairs, outs, n, mdl, score_table);
}

// -----------------------------------------------------------------------------
// affine_fill16f_kernel: the score pass FOLDED -- ONE lattice per warp, both halves of the
// registers working on it.  For batches too small to fill the GPU with one warp per PAIR of
// lattices (a shard of a fixed batch split over several GPUs: 1250 pairs are 625 warps on 592
// schedulers).  The low halves hold lattice rows [lane R, lane R + R), the high halves rows
// [32 R + lane R, ...), and the high halves run 32 columns behind the low ones: at step s lane l
// works on column s - l (low) and s - l - 32 (high).  Row 32 R - 1 (lane 31, low) therefore
// produced column j one step before row 32 R (lane 0, high) needs it, and the hand-off is the
// same single shuffle as between neighbouring lanes, rotated: lane 0 takes its HIGH top row and
// column symbol from lane 31's LOW bottom row.  Same instructions per step as
// affine_fill16u_kernel, half the rows per lane (shorter dependent chain), twice the warps.
// Exactness: as affine_fill16u_kernel (values of one LOCAL lattice in 15 bits, offset binary).
// Columns before 0 (high halves, first 32 steps) and after T (low halves, last 32) are computed
// on the "no symbol" column: they hold START-level values only (local model), never feed a real
// cell with anything a real cell would not have, and are excluded from the END bookkeeping.
template <int R>
__global__ void __launch_bounds__(32)
affine_fill16f_kernel(const AffPair *__restrict__ pairs, AffOut *__restrict__ outs, const int n,
                      const AffModel mdl, const void *__restrict__ score_table) {
    __shared__ uint32_t xt4[25];
    const int lane = threadIdx.x;
    const AffPair P = pairs[blockIdx.x];
    const int Q = P.Q, T = P.T;
    if (lane < 25) xt4[lane] = reinterpret_cast<const uint2 *>(score_table)[lane].x;  // classes 0..3, all >= 0
    __syncwarp();

    const int open = mdl.openD, one = mdl.one;
    const uint32_t openK = (uint32_t)(open * 0x10001), extDK = (uint32_t)(mdl.extD * 0x10001);
    const uint32_t extI2 = pack16(mdl.extI);
    constexpr int H = 32 * R;                 // rows per half
    const int nsteps = T + 1 + 31 + 32;
    const int rowL = lane * R, rowH = H + lane * R;

    uint32_t sel[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int il = rowL + r, ih = rowH + r;
        uint32_t sa = 0x88u, sb = 0xCCu;  // padding rows: sign of a (non-negative) pool byte = 0
        if (il >= 1 && il <= Q) { const uint32_t c = P.q[il - 1]; sa = c | ((c | 8u) << 4); }
        if (ih >= 1 && ih <= Q) { const uint32_t c = 4u + P.q[ih - 1]; sb = c | ((c | 8u) << 4); }
        sel[r] = sa | (sb << 8);
    }
    uint32_t Mp[R], Dp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        Mp[r] = kNegU16x2;
        Dp[r] = kNegU16x2;
    }
    uint32_t topM = kNegU16x2, topI = kNegU16x2, topMprev = kNegU16x2;
    uint32_t in_code = kTargetNone | (kTargetNone << 8), code0 = kTargetNone;

    uint32_t best2 = 0u;  // below every stored value
    int bjL = 0, biL = 0, bjH = 0, biH = 0;
    uint32_t pend = 0u;
    int pend_j = 0;   // column of the LOW halves; the high halves are at pend_j - 32
    auto settle_pending = [&]() {
        const uint32_t nb = __vmaxu2(best2, pend);
        if (nb != best2) {  // some half improved strictly (within a half columns arrive in increasing j)
            if ((pend & 0xFFFFu) > (best2 & 0xFFFFu)) {
                int bi = 0;
                bool found = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!found && ((Mp[r] ^ pend) & 0xFFFFu) == 0) { bi = rowL + r; found = true; }
                biL = bi;
                bjL = pend_j;
            }
            if ((pend >> 16) > (best2 >> 16)) {
                int bi = 0;
                bool found = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!found && ((Mp[r] ^ pend) >> 16) == 0) { bi = rowH + r; found = true; }
                biH = bi;
                bjH = pend_j - 32;
            }
            best2 = nb;
        }
        pend = 0u;
    };

    auto step = [&](const int s, auto ALL) {
        constexpr bool all_real = decltype(ALL)::value;   // every lane: both halves inside [0, T]
        const int j = s - lane;
        settle_pending();
        // lane 0: the low column symbol comes from the target, the high one came round from lane 31
        const uint32_t code = (lane == 0) ? (code0 | (in_code & 0xFF00u)) : in_code;
        code0 = (s + 1 <= T) ? (uint32_t)P.t[s] : (uint32_t)kTargetNone;
        const uint32_t Xa = xt4[code & 0xFFu], Xb = xt4[code >> 8];
        uint32_t cm = 0u;
        // phase A, bottom-up, rows independent: D, then G~ = max(match, D, START) + open
#pragma unroll
        for (int r = R - 1; r >= 0; --r) {
            uint32_t sc;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sc) : "r"(Xa), "r"(Xb), "r"(sel[r]));
            const uint32_t diag = (r == 0) ? topMprev : Mp[r - 1];
            Dp[r] = __vmaxu2(imad_add(Dp[r], one, extDK), Mp[r]);
            const uint32_t x = __vimax3_u16x2(imad_add(diag, one, sc), Dp[r], kBias16x2);
            Mp[r] = imad_add(x, one, openK);
        }
        // phase B, top-down: I chain (one op per row), G = max(G~, I + open) off the chain
        uint32_t Iv = __viaddmax_u16x2(topI, extI2, topM);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t gt = Mp[r];
            const uint32_t Gv = __vmaxu2(gt, imad_add(Iv, one, openK));
            Mp[r] = Gv;
            if (r & 1) cm = __vimax3_u16x2(cm, Gv, Mp[r - 1]);
            if (r + 1 < R) Iv = __viaddmax_u16x2(Iv, extI2, gt);
        }
        const uint32_t botM = Mp[R - 1], botI = Iv;
        topMprev = topM;
        if (!all_real) {   // END candidates only from real columns
            if (j < 0 || j > T) cm &= 0xFFFF0000u;
            if (j - 32 < 0 || j - 32 > T) cm &= 0x0000FFFFu;
        }
        pend = cm;
        pend_j = j;
        // rotate by one lane: lane l > 0 takes both halves from lane l - 1; lane 0 takes its HIGH
        // half from lane 31's LOW half (row 32 R follows row 32 R - 1) and has no row above its low half
        const int src = (lane + 31) & 31;
        const uint32_t nM = __shfl_sync(0xffffffffu, botM, src);
        const uint32_t nI = __shfl_sync(0xffffffffu, botI, src);
        const uint32_t nC = __shfl_sync(0xffffffffu, code, src);
        if (lane > 0) {
            topM = nM;
            topI = nI;
            in_code = nC;
        } else {
            topM = __byte_perm(nM, kNegU16x2, 0x1054);   // {low: not reachable, high: lane 31's low}
            topI = __byte_perm(nI, kNegU16x2, 0x1054);
            in_code = (nC & 0xFFu) << 8;
        }
    };

    const int fill_end = min(63, nsteps);
    const int steady_end = max(fill_end, min(T + 1, nsteps));
    int s = 0;
    for (; s < fill_end; ++s) step(s, std::false_type{});
    for (; s < steady_end; ++s) step(s, std::true_type{});
    for (; s < nsteps; ++s) step(s, std::false_type{});
    settle_pending();
    __syncwarp();

    // one lattice: the better of the two halves (max score, then min j, then min i), then over lanes
    int b = (int)(best2 & 0xFFFFu), bj = bjL, bi = biL;
    {
        const int ob = (int)(best2 >> 16);
        if ((ob > b) || (ob == b && (bjH < bj || (bjH == bj && biH < bi)))) { b = ob; bj = bjH; bi = biH; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, b, off);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if ((ob > b) || (ob == b && (oj < bj || (oj == bj && oi < bi)))) { b = ob; bj = oj; bi = oi; }
    }
    if (lane == 0) {
        AffOut o;
        o.best = b - (int)kBias16 - open;  // tracked as G = M + open, offset binary
        o.end_i = bi;
        o.end_j = bj;
        o.flags = 0;
        outs[P.out_index] = o;
    }
    (void)n;
}

// -----------------------------------------------------------------------------
// affine_fill16tb_kernel: the TRACEBACK pass with two lattices per warp.  Used for
// the banded refill of long targets and for the single pass of short ones (e.g.
// 1 kbp x 1 kbp) when every lattice of the launch qualifies.
//
// Who won is not asked with compares: every halfword is  8 * value + 1024 + TAG
// (unsigned), and the candidates of a max carry their rank in the closed-model order
// as the tag, so "first assigns, later replace only if strictly greater"
// (viterbi.c:766-775) IS the unsigned max, and the winner is read off the low bits:
//   M = max( match|3 , START|2 , D|1 , I|0 )     (T4, T5, T6, T7)
//   D = max( D<- + ext |4 , G<- |0 )              (T0 extend first, T2 open)
//   I = max( I^  + ext |4 , G^  |0 )              (T1, T3)
// Values are cleaned (& ~7) before they are used again.  Exact while
// 8 * max_sub * (min(Q,T) + 1) + 2048 < 65536 (host-checked), score' >= 0 and
// open <= ext (same vertical-chain identity as affine_fill16u_kernel; it preserves
// the tag: if G = I + open then extend beats it strictly).
// Record: one nibble per cell in the int32 kernel's layout ([sweep][step][lane][R/8]
// words per lattice), TAG FORMAT: bits 0-1 M tag, bit 2 D extended, bit 3 I extended
// (TbJob.reserved = 1 tells affine_traceback_kernel).
constexpr uint32_t kTbBias = 1024u;
constexpr uint32_t kTbNeg = (1024u - 800u) * 0x10001u;  // true -100: "not reachable"

template <int R>
__global__ void __launch_bounds__(32)
affine_fill16tb_kernel(const AffPair *__restrict__ pairs, AffOut *__restrict__ outs, const int n,
                       const AffModel mdl, const void *__restrict__ score_table) {
    constexpr int WPL = R / 8;
    __shared__ uint32_t xt4[25];
    const int lane = threadIdx.x;
    const int ia = 2 * blockIdx.x, ib = min(ia + 1, n - 1);
    const AffPair PA = pairs[ia], PB = pairs[ib];
    const bool haveB = (ia + 1 < n);
    const int QA = PA.Q, TA = PA.T, QB = PB.Q, TB_ = PB.T;
    const int T = max(TA, TB_);
    // bytes 0..3 = score' of classes 0..3 (>= 0); scaled by 8 after the PRMT (one IMAD)
    if (lane < 25) xt4[lane] = reinterpret_cast<const uint2 *>(score_table)[lane].x;
    __syncwarp();

    const int open = mdl.openD, one = mdl.one;
    const uint32_t open8 = (uint32_t)(8 * open * 0x10001);            // x + 8*open on both halves (x >= 96)
    const uint32_t extD8t = (uint32_t)((8 * mdl.extD + 4 - 1) * 0x10001);  // from Dm = D|1 to (D + ext)|4
    const uint32_t extI8t = pack16(8 * mdl.extI + 4);                  // per-half operand of VIADDMNMX.U16x2
    const uint32_t start2 = (kTbBias + 2u) * 0x10001u;                 // START: value 0, tag 2
    const int nstepsA = TA + 1 + 31, nstepsB = TB_ + 1 + 31, nsteps = T + 1 + 31;
    const int row0 = lane * R;

    uint32_t sel[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = row0 + r;
        uint32_t sa = 0x88u, sb = 0xCCu;  // padding: sign of a (non-negative) pool byte = 0
        if (i >= 1 && i <= QA) { const uint32_t c = PA.q[i - 1]; sa = c | ((c | 8u) << 4); }
        if (i >= 1 && i <= QB) { const uint32_t c = 4u + PB.q[i - 1]; sb = c | ((c | 8u) << 4); }
        sel[r] = sa | (sb << 8);
    }
    uint32_t Gp[R], Dm[R];  // clean G = M + open, and D|1, of the previous column
#pragma unroll
    for (int r = 0; r < R; ++r) {
        Gp[r] = kTbNeg;
        Dm[r] = kTbNeg | 0x00010001u;
    }
    uint32_t topG = kTbNeg, topI = kTbNeg, topGprev = kTbNeg;
    uint32_t in_code = kTargetNone | (kTargetNone << 8), code0 = in_code;
    uint32_t *tbA = PA.tb + (size_t)lane * WPL, *tbB = PB.tb + (size_t)lane * WPL;

    uint32_t best2 = 0u, pend = 0u;
    int bjA = 0, biA = 0, bjB = 0, biB = 0, pend_j = 0;
    auto settle_pending = [&]() {
        const uint32_t nb = __vmaxu2(best2, pend);
        if (nb != best2) {
            if ((pend & 0xFFFFu) > (best2 & 0xFFFFu)) {
                int bi = 0;
                bool found = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!found && ((Gp[r] ^ pend) & 0xFFFFu) == 0) { bi = row0 + r; found = true; }
                biA = bi;
                bjA = pend_j;
            }
            if ((pend >> 16) > (best2 >> 16)) {
                int bi = 0;
                bool found = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!found && ((Gp[r] ^ pend) >> 16) == 0) { bi = row0 + r; found = true; }
                biB = bi;
                bjB = pend_j;
            }
            best2 = nb;
        }
        pend = 0u;
    };

    for (int s = 0; s < nsteps; ++s) {
        const int j = s - lane;
        settle_pending();
        const uint32_t code = (lane == 0) ? code0 : in_code;
        {
            const uint32_t ca = (s + 1 <= TA) ? (uint32_t)PA.t[s] : (uint32_t)kTargetNone;
            const uint32_t cb = (s + 1 <= TB_) ? (uint32_t)PB.t[s] : (uint32_t)kTargetNone;
            code0 = ca | (cb << 8);
        }
        uint32_t botG = kTbNeg, botI = kTbNeg;
        if (j >= 0 && j <= T) {
            const uint32_t Xa = xt4[code & 0xFFu], Xb = xt4[code >> 8];
            uint32_t acc[R / 4];
#pragma unroll
            for (int k = 0; k < R / 4; ++k) acc[k] = 0u;
            uint32_t Gt[R];  // G~ = max(match, START, D) + open, clean
            // phase A (rows independent; bottom-up, so that row r still finds the previous
            // column's G of row r-1 in place): D, then G~
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                uint32_t sc;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sc) : "r"(Xa), "r"(Xb), "r"(sel[r]));
                const uint32_t diag = (r == 0) ? topGprev : Gp[r - 1];
                // D: extend (tag 4) first, open (tag 0) replaces only if strictly greater
                const uint32_t Dt = __vmaxu2(imad_add(Dm[r], one, extD8t), Gp[r]);
                const uint32_t dm = (Dt & 0xFFF8FFF8u) | 0x00010001u;          // clean, M-level tag 1
                // match candidate: 8 * (G_diag + score') + tag 3
                const uint32_t x = imad_add(sc, 8, diag) + 0x00030003u;
                const uint32_t mt = __vimax3_u16x2(x, start2, dm);            // T4, T5, T6
                Gt[r] = imad_add(mt & 0xFFF8FFF8u, one, open8);
                // record: M tag (bits 0-1, completed in phase B) and D's bit 2
                acc[r / 4] |= ((mt & 0x00030003u) | (Dt & 0x00040004u)) << (4 * (r % 4));
                Dm[r] = dm;
                Gp[r] = mt;  // tagged max without I; phase B finishes it
            }
            // phase B: I chain (one dependent op per row); M = max(mt, I|0)
            uint32_t It = __viaddmax_u16x2(topI, extI8t, topG);   // I of row 0: extend|4 vs open|0
            uint32_t cm = 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t ic = It & 0xFFF8FFF8u;                          // clean, M-level tag 0
                const uint32_t mt = Gp[r];
                // if I wins (strictly greater than every earlier candidate) the tag becomes 0
                const uint32_t m2 = __vmaxu2(mt, ic);
                const uint32_t Gv = __vmaxu2(Gt[r], imad_add(ic, one, open8));  // clean G of this cell
                // fix the M tag where I won: (m2 & 3) replaces (mt & 3); plus I's bit 3
                acc[r / 4] ^= (((mt ^ m2) & 0x00030003u) | ((It & 0x00040004u) << 1)) << (4 * (r % 4));
                Gp[r] = Gv;
                if (r & 1) cm = __vimax3_u16x2(cm, Gv, Gp[r - 1]);
                if (r + 1 < R) It = __viaddmax_u16x2(ic, extI8t, Gt[r]);
                else botI = ic;
            }
            botG = Gp[R - 1];
            topGprev = topG;
            pend = cm;
            pend_j = j;
            // nibbles: acc[k] = rows 4k..4k+3, lattice A in the low half, B in the high half
            if (s < nstepsA && j <= TA) {
#pragma unroll
                for (int w = 0; w < WPL; ++w) tbA[w] = __byte_perm(acc[2 * w], acc[2 * w + 1], 0x5410);
            }
            if (haveB && s < nstepsB && j <= TB_) {
#pragma unroll
                for (int w = 0; w < WPL; ++w) tbB[w] = __byte_perm(acc[2 * w], acc[2 * w + 1], 0x7632);
            }
        }
        tbA += 32 * WPL;
        tbB += 32 * WPL;
        const uint32_t nG = __shfl_up_sync(0xffffffffu, botG, 1);
        const uint32_t nI = __shfl_up_sync(0xffffffffu, botI, 1);
        const uint32_t nC = __shfl_up_sync(0xffffffffu, code, 1);
        if (lane > 0) {
            topG = nG;
            topI = nI;
            in_code = nC;
        }
    }
    settle_pending();
    __syncwarp();

    int bA = (int)(best2 & 0xFFFFu), bB = (int)(best2 >> 16);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bA, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bjA, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biA, off);
            if ((ob > bA) || (ob == bA && (oj < bjA || (oj == bjA && oi < biA)))) { bA = ob; bjA = oj; biA = oi; }
        }
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bB, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bjB, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biB, off);
            if ((ob > bB) || (ob == bB && (oj < bjB || (oj == bjB && oi < biB)))) { bB = ob; bjB = oj; biB = oi; }
        }
    }
    if (lane == 0) {
        AffOut o;
        o.best = (bA - (int)kTbBias) / 8 - open;  // tracked as clean 8 * (M + open) + bias
        o.end_i = biA;
        o.end_j = bjA;
        o.flags = 0;
        outs[PA.out_index] = o;
        if (haveB) {
            o.best = (bB - (int)kTbBias) / 8 - open;
            o.end_i = biB;
            o.end_j = bjB;
            outs[PB.out_index] = o;
        }
    }
}

}  // namespace c4b
