// e2g_systolic.cuh -- est2genome lattice fill + traceback (spliced cDNA vs genome).
//
// Hand-specialised counterpart of the generated optimal:est2genome DP functions
// (src/c4/viterbi.c:1638-1727 over the model of src/model/est2genome.c:57-93 +
// src/model/intron.c:588-697), for the closed model printed in SURVEY.md §8a
// (10 states, 24 transitions, one shadow slot, max advance (1,2)); the template
// is verified transition by transition on the host (analyze_est2genome).
// Per strand x in {forward, reverse} and cell (i,j), candidates in CLOSED-MODEL
// ORDER (first assigns, later replace only if strictly greater, viterbi.c:766-775):
//   N_x (intron): open  M_x(i,j-2) + intron_open + splice_pre[j-2]      (T3 / T0)
//                 loop  N_x(i,j-1)                                       (T4 / T1)
//   I_x:          open  M_x(i-1,j) + gap_open   then  I_x(i-1,j) + ext   (T12,T14 / T7,T9)
//   D_x:          open  M_x(i,j-1) + gap_open   then  D_x(i,j-1) + ext   (T13,T15 / T8,T10)
//   M_x:          close N_x(i,j-2) + splice_post[j-2] if min <= len <= max else -inf,
//                       UNDERFLOW-clamped (intron.c:151-159)             (T5 / T2)
//                 match M_x(i-1,j-1) + s                                 (T11 / T6)
//                 I_x(i,j), D_x(i,j), START 0                            (T19,T20,T21 / T16,T17,T18)
//   END:          M_reverse first, M_forward only if strictly greater    (T22, T23)
// The shadow (intron start, viterbi.c:413-422) is the column where N was opened; it
// rides along N only (no other state's copy is ever read).
//
// Mapping: one CTA per lattice, W warps = W strips of 256 lattice rows (8 rows per
// lane in registers, 14 loop-carried values per row), every warp a skewed systolic
// wavefront as in affine_systolic.cuh; the strips run concurrently as a pipeline:
// strip w hands its bottom row to strip w+1 through a shared-memory ring with
// producer/consumer counters (no __syncthreads in the fill).  The traceback record
// is 13 bit/cell (2 x {3-bit M winner, N, I, D} + END strand) stored as one
// halfword per cell, 16 B per lane per step, coalesced in production order.
#pragma once
#include <type_traits>

#include "c4b_common.cuh"

namespace c4b {

constexpr int kE2gR = 8;          // rows per lane
constexpr int kE2gRing = 128;     // hand-off ring depth (columns)
constexpr int kE2gMaxWarps = 8;   // <= 2048 lattice rows

struct E2gPair {
    const uint8_t *q;       // query classes (PRMT) per position
    const uint8_t *t;       // target column codes per position
    const uint32_t *sp;     // per target position: int8 x4 {ss5_fwd, ss3_fwd, ss5_rev, ss3_rev}
    int32_t Q, T;
    uint16_t *tb;           // [warp][step][lane][8] halfwords, or null
    int64_t out_index;
};

struct E2gOut {
    int32_t best, end_i, end_j, end_forward;
};

struct E2gModel {
    int32_t open, ext, intron_open, min_intron, max_intron, one;
    // transition ids by role, forward then reverse strand
    int32_t tNopen[2], tNloop[2], tNclose[2], tMatch[2], tIopen[2], tDopen[2], tIext[2], tDext[2];
    int32_t tI2M[2], tD2M[2], tS2M[2], tM2E[2];
    int32_t win_cols;   // packed kernel, windowed traceback: columns per checkpoint window (a power of two)
};

__device__ __forceinline__ int e2g_prmt_sx(uint32_t lo, uint32_t hi, uint32_t sel) {
    int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
}

// one strand of one cell, phase A: everything that only needs previous columns
struct E2gStrandRow {
    int G1, G2, D1, N1, N2, S1, S2;  // loop-carried (G = M + gap_open)
};

template <bool TB>
__global__ void __launch_bounds__(32 * kE2gMaxWarps)
e2g_fill_kernel(const E2gPair *__restrict__ pairs, E2gOut *__restrict__ outs, const E2gModel mdl,
                const uint2 *__restrict__ score_table) {
    constexpr int R = kE2gR;
    __shared__ uint2 xtab[25];
    __shared__ int4 ring[kE2gMaxWarps][kE2gRing];
    __shared__ volatile int prod[kE2gMaxWarps], cons[kE2gMaxWarps];
    __shared__ int red[kE2gMaxWarps][4];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const E2gPair P = pairs[blockIdx.x];
    const int Q = P.Q, T = P.T;
    if (threadIdx.x < 25) xtab[threadIdx.x] = score_table[threadIdx.x];
    if (threadIdx.x < kE2gMaxWarps) { prod[threadIdx.x] = 0; cons[threadIdx.x] = 0; }
    __syncthreads();

    const int open = mdl.open, ext = mdl.ext;
    const int min_intron = mdl.min_intron, max_intron = mdl.max_intron;
    const int nsteps = T + 1 + 31;
    const int row0 = warp * 32 * R + lane * R;
    const bool first_row_lane = (warp == 0 && lane == 0);
    const bool consumer = (warp > 0), producer = (warp + 1 < nwarps);

    uint32_t sel[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = row0 + r;
        const int c = (i >= 1 && i <= Q) ? P.q[i - 1] : kPadClass;
        sel[r] = (uint32_t)c * 0x1111u | 0x8880u;
    }
    E2gStrandRow st[2][R];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            st[x][r].G1 = st[x][r].G2 = st[x][r].D1 = st[x][r].N1 = st[x][r].N2 = NEG2;
            st[x][r].S1 = st[x][r].S2 = 0;
        }
    // row above my strip, this column and the previous one (for the diagonal)
    int topG[2] = {NEG2, NEG2}, topI[2] = {NEG2, NEG2}, topGprev[2] = {NEG2, NEG2};
    int in_code = kTargetNone, code0 = kTargetNone;
    uint32_t in_sp = 0, sp0 = 0;  // splice word of column j-2 handed down the lanes
    uint16_t *tbp = nullptr;
    if (TB) tbp = P.tb + (((size_t)warp * nsteps) * 32 + lane) * R;
    int best = INT32_MIN, best_i = 0, best_j = 0, best_f = 0;

    for (int s = 0; s < nsteps; ++s) {
        const int j = s - lane;
        // ---- column inputs for lane 0: symbol of column s, splice scores of column s-2
        int code = (lane == 0) ? code0 : in_code;
        uint32_t spw = (lane == 0) ? sp0 : in_sp;
        code0 = (s + 1 <= T) ? (int)P.t[s] : kTargetNone;
        sp0 = (s + 1 >= 2 && s + 1 <= T) ? P.sp[s - 1] : 0u;   // source column (s+1)-2
        if (consumer) {
            // wait for the strip above to publish column s (its lane 31 is 31 steps behind)
            if (s <= T) {
                while (prod[warp - 1] <= s) __nanosleep(20);
                __threadfence_block();  // the column was written before the counter moved
                const int4 v = ring[warp - 1][s & (kE2gRing - 1)];
                if (lane == 0) { topG[0] = v.x; topI[0] = v.y; topG[1] = v.z; topI[1] = v.w; }
                __syncwarp();
                if (lane == 0) cons[warp] = s + 1;
            }
        }
        int botG[2] = {NEG2, NEG2}, botI[2] = {NEG2, NEG2};
        if (j >= 0 && j <= T) {
            const uint2 X = xtab[code];
            // per-column intron terms (source column j-2); meaningless (and unused:
            // the N/M inputs are sentinels) while j < 2
            const int b0 = (int)(int8_t)(spw & 255u), b1 = (int)(int8_t)((spw >> 8) & 255u);
            const int b2 = (int)(int8_t)((spw >> 16) & 255u), b3 = (int)(int8_t)(spw >> 24);
            // forward: open at a 5' site, close at a 3' site; reverse: 3' then 5'
            const int pre[2] = {mdl.intron_open + b0 - open, mdl.intron_open + b3 - open};  // added to G2
            const int post[2] = {b1, b2};
            const int jsrc = j - 2;  // region-relative source column = intron start stamp
            int Xc[2][R];             // running first-max of the M candidates (T5, T11)
            uint32_t bits[R];         // traceback halfword per row
            int Dn[2][R];
            // ---- phase A: everything that depends on previous columns only ----------
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int sc = e2g_prmt_sx(X.x, X.y, sel[r]);
                uint32_t b = 0;
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    E2gStrandRow &S = st[x][r];
                    // N: open (T3/T0) first, loop (T4/T1) replaces if strictly greater
                    int nv = max(S.G2 + pre[x], LOW);     // UNDERFLOW protect of the splice calc
                    int ns = jsrc;
                    const bool loop = nv < S.N1;
                    if (loop) { nv = S.N1; ns = S.S1; }
                    // M candidate T5/T2: close an intron opened at column S2
                    const int len = jsrc - S.S2 + 2;
                    int c5 = (len >= min_intron && len <= max_intron) ? S.N2 + post[x] : LOW;
                    c5 = max(c5, LOW);
                    // M candidate T11/T6: match
                    const int diag = (r == 0) ? topGprev[x] : st[x][r - 1].G1;
                    const int c11 = diag + sc;
                    int xc = c5, dm = 0;
                    if (xc < c11) { xc = c11; dm = 1; }
                    // D: open (T13/T8) first, extend (T15/T10) if strictly greater
                    int dv = S.G1;
                    const int de = S.D1 + ext;
                    const bool dext = dv < de;
                    if (dext) dv = de;
                    Xc[x][r] = xc;
                    Dn[x][r] = dv;
                    // shift the intron history now (G shifts in phase B)
                    S.N2 = S.N1; S.S2 = S.S1; S.N1 = nv; S.S1 = ns; S.D1 = dv;
                    if (TB) b |= ((uint32_t)dm | ((uint32_t)loop << 3) | ((uint32_t)dext << 5)) << (6 * x);
                }
                bits[r] = b;
            }
            // ---- phase B: the vertical chain I -> M -> G, top-down ---------------------
            int upG[2] = {topG[0], topG[1]}, upI[2] = {topI[0], topI[1]};
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int mval[2];
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    E2gStrandRow &S = st[x][r];
                    // I: open (T12/T7) first, extend (T14/T9) if strictly greater
                    int iv = upG[x];
                    const int ie = upI[x] + ext;
                    const bool iext = iv < ie;
                    if (iext) iv = ie;
                    // M: (T5,T11 in Xc) then I (T19/T16), D (T20/T17), START (T21/T18)
                    int m = Xc[x][r];
                    uint32_t dm = (bits[r] >> (6 * x)) & 1u;
                    if (m < iv) { m = iv; dm = 2; }
                    if (m < Dn[x][r]) { m = Dn[x][r]; dm = 3; }
                    if (m < 0) { m = 0; dm = 4; }
                    const int g = m + open;
                    S.G2 = S.G1;
                    S.G1 = g;
                    upG[x] = g;
                    upI[x] = iv;
                    mval[x] = m;
                    if (TB) bits[r] = (bits[r] & ~(7u << (6 * x))) | ((dm | ((uint32_t)iext << 4)) << (6 * x));
                }
                // END: reverse strand first (T22), forward (T23) only if strictly greater
                const bool endf = mval[1] < mval[0];
                const int e = endf ? mval[0] : mval[1];
                if (TB) bits[r] |= (uint32_t)endf << 12;
                const int i = row0 + r;
                if (i <= Q && (e > best || (e == best && j < best_j))) {
                    best = e; best_i = i; best_j = j; best_f = endf;
                }
            }
            botG[0] = upG[0]; botI[0] = upI[0]; botG[1] = upG[1]; botI[1] = upI[1];
            topGprev[0] = topG[0];
            topGprev[1] = topG[1];
            if (TB) {
                uint4 w;
                w.x = bits[0] | (bits[1] << 16); w.y = bits[2] | (bits[3] << 16);
                w.z = bits[4] | (bits[5] << 16); w.w = bits[6] | (bits[7] << 16);
                *reinterpret_cast<uint4 *>(tbp) = w;
            }
            if (producer && lane == 31) {
                // do not overrun the consumer, then publish column j
                while (j - cons[warp + 1] >= kE2gRing) __nanosleep(20);
                ring[warp][j & (kE2gRing - 1)] = make_int4(botG[0], botI[0], botG[1], botI[1]);
                __threadfence_block();
                prod[warp] = j + 1;
            }
        }
        if (first_row_lane) { topG[0] = topG[1] = NEG2; topI[0] = topI[1] = NEG2; }
        if (TB) tbp += 32 * R;
        // hand-off to the next lane
        const int n0 = __shfl_up_sync(0xffffffffu, botG[0], 1), n1 = __shfl_up_sync(0xffffffffu, botI[0], 1);
        const int n2 = __shfl_up_sync(0xffffffffu, botG[1], 1), n3 = __shfl_up_sync(0xffffffffu, botI[1], 1);
        const int nC = __shfl_up_sync(0xffffffffu, code, 1);
        const uint32_t nS = __shfl_up_sync(0xffffffffu, spw, 1);
        if (lane > 0) {
            topG[0] = n0; topI[0] = n1; topG[1] = n2; topI[1] = n3;
            in_code = nC;
            in_sp = nS;
        }
    }
    // ---- lexicographic reduction (score, -j, -i) over the CTA -------------------------
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, off), oj = __shfl_xor_sync(0xffffffffu, best_j, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, off), of = __shfl_xor_sync(0xffffffffu, best_f, off);
        if (ob > best || (ob == best && (oj < best_j || (oj == best_j && oi < best_i)))) {
            best = ob; best_j = oj; best_i = oi; best_f = of;
        }
    }
    if (lane == 0) { red[warp][0] = best; red[warp][1] = best_j; red[warp][2] = best_i; red[warp][3] = best_f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nwarps; ++w) {
            const int ob = red[w][0], oj = red[w][1], oi = red[w][2];
            if (ob > best || (ob == best && (oj < best_j || (oj == best_j && oi < best_i)))) {
                best = ob; best_j = oj; best_i = oi; best_f = red[w][3];
            }
        }
        E2gOut o;
        o.best = best; o.end_i = best_i; o.end_j = best_j; o.end_forward = best_f;
        outs[P.out_index] = o;
    }
}

struct E2gJob {
    int32_t pair, result, q_origin, t_origin;
    int64_t ops_off;
    int32_t ops_cap, reserved;
};

// Viterbi_Data_create_Alignment (viterbi.c:342-392) over the 13-bit records.
__global__ void e2g_traceback_kernel(const E2gPair *__restrict__ pairs, const E2gOut *__restrict__ outs,
                                     const E2gJob *__restrict__ jobs, int n, const E2gModel mdl, int threshold,
                                     c4b_result *__restrict__ results, int32_t *__restrict__ ops) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const E2gJob J = jobs[g];
    const E2gPair P = pairs[J.pair];
    const E2gOut o = outs[P.out_index];
    const int nsteps = P.T + 1 + 31;
    c4b_result res;
    res.score = o.best; res.status = 0; res.reserved = 0; res.n_ops = 0; res.ops_offset = J.ops_off;
    int i = o.end_i, j = o.end_j;
    res.query_end = J.q_origin + i;
    res.target_end = J.t_origin + j;
    int32_t *out = ops + 2 * J.ops_off;
    int n_runs = 0, last_t = -1;
    bool overflow = false;
    auto emit = [&](int t) {
        if (t == last_t) out[2 * (n_runs - 1) + 1] += 1;
        else if (n_runs < J.ops_cap) { out[2 * n_runs] = t; out[2 * n_runs + 1] = 1; ++n_runs; last_t = t; }
        else overflow = true;
    };
    auto record = [&](int ci, int cj) -> uint32_t {
        const int w = ci / (32 * kE2gR), ln = (ci / kE2gR) & 31, r = ci % kE2gR;
        return P.tb[(((size_t)w * nsteps + (cj + ln)) * 32 + ln) * kE2gR + r];
    };
    if (o.best < threshold) {
        res.status = 1;
    } else {
        // x: 0 = forward strand fields (bits 0..5), 1 = reverse (bits 6..11)
        int x = (record(i, j) >> 12) & 1u ? 0 : 1;
        int state = 0;  // 0 M, 1 I, 2 D, 3 N
        emit(mdl.tM2E[x]);
        for (;;) {
            const uint32_t f = (record(i, j) >> (6 * x)) & 63u;
            if (state == 0) {
                const uint32_t dm = f & 7u;
                if (dm == 0) { emit(mdl.tNclose[x]); j -= 2; state = 3; }
                else if (dm == 1) { emit(mdl.tMatch[x]); --i; --j; }
                else if (dm == 2) { emit(mdl.tI2M[x]); state = 1; }
                else if (dm == 3) { emit(mdl.tD2M[x]); state = 2; }
                else { emit(mdl.tS2M[x]); break; }
            } else if (state == 1) {
                if (f & 16u) emit(mdl.tIext[x]); else { emit(mdl.tIopen[x]); state = 0; }
                --i;
            } else if (state == 2) {
                if (f & 32u) emit(mdl.tDext[x]); else { emit(mdl.tDopen[x]); state = 0; }
                --j;
            } else {
                if (f & 8u) { emit(mdl.tNloop[x]); --j; }
                else { emit(mdl.tNopen[x]); j -= 2; state = 0; }
            }
            if (i < 0 || j < 0 || overflow) { res.status = 4; break; }
        }
        for (int a = 0, b = n_runs - 1; a < b; ++a, --b) {
            const int t0 = out[2 * a], l0 = out[2 * a + 1];
            out[2 * a] = out[2 * b]; out[2 * a + 1] = out[2 * b + 1];
            out[2 * b] = t0; out[2 * b + 1] = l0;
        }
    }
    res.n_ops = (res.status == 0) ? n_runs : 0;
    res.query_start = J.q_origin + max(i, 0);
    res.target_start = J.t_origin + max(j, 0);
    results[J.result] = res;
}

__global__ void e2g_score_results_kernel(const E2gPair *__restrict__ pairs, const E2gOut *__restrict__ outs,
                                         const int32_t *__restrict__ q_origin,
                                         const int32_t *__restrict__ t_origin, int n,
                                         c4b_result *__restrict__ results) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const E2gOut o = outs[pairs[p].out_index];
    c4b_result r;
    r.score = o.best;
    r.query_start = q_origin[p]; r.target_start = t_origin[p];
    r.query_end = q_origin[p] + o.end_i; r.target_end = t_origin[p] + o.end_j;
    r.n_ops = 0; r.ops_offset = 0; r.status = 0; r.reserved = 0;
    results[pairs[p].out_index] = r;
}

}  // namespace c4b
