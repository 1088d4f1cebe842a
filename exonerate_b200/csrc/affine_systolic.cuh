// affine_systolic.cuh -- the affine-gap lattice fill (hot kernel of the path).
//
// Replaces the generated Viterbi_DP_Func of the affine family
// (optimal:affine:{local,global,bestfit,overlap}:* find score / find path,
// generator src/c4/viterbi.c:1638-1727, semantics src/c4/viterbi.c:655-837)
// for models whose closed form is the 5-state / 9-transition template printed
// in SURVEY.md §8a:
//   T0 D->D (0,1) ext   T1 I->I (1,0) ext   T2 M->D (0,1) open  T3 M->I (1,0) open
//   T4 M->M (1,1) match T5 S->M (0,0)       T6 D->M (0,0)       T7 I->M (0,0)
//   T8 M->E (0,0)
// "first valid transition assigns, later ones replace only if strictly greater"
// (viterbi.c:766-775) in exactly this order is what every max below encodes.
//
// Mapping (B200): one CTA = one warp = one query x target lattice.  Lane l owns
// R consecutive lattice rows in REGISTERS; lanes are skewed by one column
// (lane l works on column step-l), so the anti-diagonal wavefront is the warp
// itself: the bottom cell of a lane's strip reaches the next lane through ONE
// warp shuffle per value per step, the target symbol rides the same shuffle,
// and no cell value ever touches shared memory or HBM.  Queries longer than
// 32*R-1 are swept in strips; the strip hand-off row {G,I}[T+1] lives in L2.
// int32 max-plus on the DPX pipe (VIADDMNMX / VIMNMX3 / VIMNMX.RELU); the
// substitution score is one PRMT on a byte-packed column of the matrix.
//
// TB=true additionally records, per cell, which transition won for M (2 bits),
// D (1 bit) and I (1 bit): 4 bits/cell, R/2 bytes per lane per step, written as
// one coalesced vector store in the skewed order the warp produces them.  The
// winners are not found with compares: TB works on 8 * value + rank tag (see the
// kernel), nibble = {bits 0-1 rank of M's winner (3 match .. 0 insert), bit 2 D
// extended, bit 3 I extended} -- the same format as affine_fill16tb_kernel.
#pragma once
#include <type_traits>

#include "c4b_common.cuh"

namespace c4b {

enum { END_ANYWHERE = 0, END_RESTRICTED = 1 };
enum { SCORE_PRMT = 0, SCORE_SMEM = 1 };

// prmt.b32 with the full 4-bit selector (bit 3 = replicate the sign of the
// selected byte); __byte_perm masks the selector to 3 bits and cannot do this.
__device__ __forceinline__ int prmt_sx(uint32_t lo, uint32_t hi, uint32_t sel) {
    int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
}

// M + gap_open.  Written as a multiply-add by a run-time 1 so that it issues on
// the FMA pipe (IMAD) and leaves the ALU/DPX pipe, the binding one, to the
// max-plus instructions (profiles/: pipe_alu 86 %, pipe_fma 12 %).
__device__ __forceinline__ int add_open(int m, int one, int open) {
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(m), "r"(one), "r"(open));
    return d;
}

// score_table: SCORE_PRMT -> uint2[25]  (bytes k=0..7 = s(class k, column code) - open)
//              SCORE_SMEM -> int32[25*25] (row 24 = pad rows, column 24 = "no symbol")
// ENDMODE END_ANYWHERE is only instantiated for START=END=ANYWHERE (local) models.
// blockDim.x = 32 W: the W warps of a CTA take the sweeps (strips of 32 R rows) of ONE
// lattice round-robin and run them concurrently as a pipeline -- sweep k+1 follows sweep
// k at a distance of >= 32 columns, reading the hand-off row {G, I} the moment it is
// published (a monotone counter in shared memory, no __syncthreads in the fill).  A
// long query therefore occupies up to 8 schedulers instead of one.  W = 1 for batches
// whose queries fit one sweep.
constexpr int kAffMaxWarps = 8;

// BLK = true: some lattice of the launch carries SubOpt blocked cells (src/c4/subopt.h:77-80):
// at a blocked DESTINATION cell the MATCH-labelled transition (T4) is skipped
// (viterbi.c:701-704).  The host turns each lattice's list into {column, row mask} entries
// per lane strip, sorted by column (AffPair::blk / blk_off); a lane walks its strip's entries
// with one cursor, and only the steps on which some lane meets a blocked column run the
// masked copy of phase A (the diagonal candidate becomes "not reachable").
template <int R, bool TB, int ENDMODE, int SM, bool BLK = false>
__global__ void __launch_bounds__(32 * kAffMaxWarps)
affine_fill_kernel(const AffPair *__restrict__ pairs, AffOut *__restrict__ outs,
                   const AffModel mdl, const void *__restrict__ score_table) {
    constexpr int WPL = R / 8;  // traceback words per lane per step
    constexpr bool LOCAL = (ENDMODE == END_ANYWHERE);
    __shared__ uint2 xtab[25];
    __shared__ int32_t subm[SM == SCORE_SMEM ? 25 * 25 : 1];

    // published columns of the sweep each warp is (or was last) working on, as a
    // monotone count: sweep * (T + 1) + columns whose bottom row is in the hand-off array
    __shared__ volatile long long vprog[kAffMaxWarps];
    __shared__ int red[kAffMaxWarps][3];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const AffPair P = pairs[blockIdx.x];
    const int Q = P.Q, T = P.T;

    if (SM == SCORE_PRMT) {
        if (threadIdx.x < 25) xtab[threadIdx.x] = reinterpret_cast<const uint2 *>(score_table)[threadIdx.x];
    } else {
        for (int k = threadIdx.x; k < 25 * 25; k += blockDim.x)
            subm[k] = reinterpret_cast<const int32_t *>(score_table)[k];
    }
    if (threadIdx.x < kAffMaxWarps) vprog[threadIdx.x] = 0;
    __syncthreads();

    const int open = mdl.openD, extD = mdl.extD, extI = mdl.extI;  // openD == openI (checked)
    const int one = mdl.one;
    // TB = true works on 8 * value + TAG: the candidates of a max carry their rank in the
    // closed-model order in the low bits, so "first assigns, later replace only if strictly
    // greater" IS the max and the winner is read off the tag (cleaned with & ~7 before reuse):
    //   M = max(match|3, START|2, D|1, I|0)   D = max(D<- + ext |4, G<- |0)   I likewise
    constexpr int NEGK = TB ? -(1 << 28) : NEG2;   // "not reachable" in the kernel's units
    const int open8 = 8 * open, extD8t = 8 * extD + 4 - 1 /* from D|1 */, extI8t = 8 * extI + 4;
    // where may START be entered / END be left (src/c4/layout.c:21-88)
    const int ss = mdl.start_scope, es = mdl.end_scope;
    const bool start_any = (ss == C4B_SCOPE_ANYWHERE);
    const bool start_row0 = start_any || ss == C4B_SCOPE_EDGE || ss == C4B_SCOPE_QUERY;
    const bool start_col0 = start_any || ss == C4B_SCOPE_EDGE || ss == C4B_SCOPE_TARGET;
    const bool end_rowQ = (es == C4B_SCOPE_EDGE || es == C4B_SCOPE_QUERY);
    const bool end_colT = (es == C4B_SCOPE_EDGE || es == C4B_SCOPE_TARGET);
    const bool corner_only = (es == C4B_SCOPE_CORNER);

    const int rows_per_sweep = 32 * R;
    const int nsweeps = (Q + 1 + rows_per_sweep - 1) / rows_per_sweep;
    const int nsteps = T + 1 + 31;
    const int pub_mask = (T >= 4096) ? 15 : 3;   // hand-off rows are published every 16 (4) columns

    // first strict maximum of END in (target outer, query inner) order
    // (viterbi.c:778-791) == lexicographic max of (score, -j, -i); tracked on
    // G = M + open and converted at the end
    int best = INT32_MIN, best_j = 0, best_i = 0;

    for (int sweep = warp; sweep < nsweeps; sweep += W) {
        const int row0 = sweep * rows_per_sweep + lane * R;  // lattice row of r = 0
        const bool first_row_lane = (sweep == 0 && lane == 0);
        const bool later_sweep = (sweep > 0);
        // per-row query operand: PRMT selector or matrix row offset
        uint32_t sel[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = row0 + r;  // lattice row; consumes query symbol i-1
            if (SM == SCORE_PRMT) {
                const int c = (i >= 1 && i <= Q) ? P.q[i - 1] : kPadClass;
                sel[r] = (uint32_t)c * 0x1111u | 0x8880u;  // byte c, sign-extended
            } else {
                const int c = (i >= 1 && i <= Q) ? P.q[i - 1] : 24;
                sel[r] = (uint32_t)c * 25u;
            }
        }
        // loop-carried state: G = M + open and D of the previous column, per row
        int Mp[R], Dp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            Mp[r] = NEGK;
            Dp[r] = TB ? (NEGK | 1) : NEG2;   // TB: D is kept with its M-level tag
        }
        // END at the last lattice row: which of my registers holds row Q
        const int rQ = (Q >= row0 && Q < row0 + R) ? (Q - row0) : -1;

        const int2 *top_in = (sweep & 1) ? P.top0 : P.top1;  // written by sweep-1
        int2 *top_out = (sweep & 1) ? P.top1 : P.top0;
        const bool write_top = (sweep + 1 < nsweeps) && (lane == 31);

        int topM = NEGK, topI = NEGK, topMprev = NEGK;  // row above my strip
        int in_code = kTargetNone;                      // column code handed down
        int code0 = kTargetNone;                        // lane 0: column 0 has no symbol
        // hand-off from the sweep above, which another warp may still be producing
        const bool piped = later_sweep && W > 1;
        const int wp = (sweep - 1) % W;
        const long long in_base = (long long)(sweep - 1) * (T + 1);
        long long avail = 0;   // last value seen of the producer's counter
        auto wait_column = [&](int col) {   // until column `col` of the sweep above is published
            const long long need = in_base + col + 1;
            if (avail < need) {
                while ((avail = vprog[wp]) < need) __nanosleep(40);
                __threadfence_block();
            }
        };
        int2 top0v = make_int2(NEGK, NEGK);
        if (later_sweep) {
            if (piped) wait_column(0);
            top0v = __ldcg(top_in);                     // uniform load, lane 0 uses it
        }
        uint32_t *tbp = nullptr;
        if (TB) tbp = P.tb + ((size_t)sweep * nsteps * 32 + lane) * WPL;
        // SubOpt: cursor into my strip's blocked columns (columns are relative to P.blk_j0,
        // the band origin of a traceback refill)
        int blk_cur = 0, blk_end = 0, blk_next = INT32_MAX;
        if (BLK && P.blk_off) {
            blk_cur = P.blk_off[sweep * 32 + lane];
            blk_end = P.blk_off[sweep * 32 + lane + 1];
            int lo = blk_cur, hi = blk_end;      // first entry at or after the band origin
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (P.blk[mid].x < P.blk_j0) lo = mid + 1; else hi = mid;
            }
            blk_cur = lo;
            if (blk_cur < blk_end) blk_next = P.blk[blk_cur].x - P.blk_j0;
        }

        // END_ANYWHERE: the column maximum of a step is examined at the top of the
        // NEXT step, when the column's values sit in their loop-carried registers
        // (checking right after the row loop made ptxas copy all 2R registers on
        // the common no-update path, +1 instruction per cell).
        int pend_cm = INT32_MIN, pend_j = 0;
        auto settle_pending = [&]() {
            if (pend_cm > best || (pend_cm == best && pend_j < best_j)) {
                int bi = 0;
                bool found = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!found && Mp[r] == pend_cm) { bi = row0 + r; found = true; }
                best = pend_cm;
                best_j = pend_j;
                best_i = bi;
            }
            pend_cm = INT32_MIN;
        };

        // One step of the wavefront.  ALL = std::true_type in the steady state, where
        // every lane has a column in [0, T] and the activity test (and its divergence
        // bookkeeping) disappears from the instruction stream.
        auto step = [&](const int s, auto ALL) {
            constexpr bool all_active = decltype(ALL)::value;
            const int j = s - lane;
            if (LOCAL) settle_pending();
            // the column symbol / hand-off row are fetched with warp-uniform addresses
            // (one broadcast transaction, no divergent branch); only lane 0 consumes them
            int code = (lane == 0) ? code0 : in_code;
            if (later_sweep && lane == 0) {
                topM = top0v.x;
                topI = top0v.y;
            }
            if (s + 1 <= T) {
                code0 = (int)P.t[s];
                if (later_sweep) {
                    if (piped) wait_column(s + 1);
                    top0v = __ldcg(top_in + s + 1);
                }
            } else {
                code0 = kTargetNone;
            }
            int botM = NEGK, botI = NEGK;
            uint32_t bmask = 0;   // rows of my strip whose cell in this column is blocked
            if (BLK && j == blk_next) {
                bmask = (uint32_t)P.blk[blk_cur].y;
                ++blk_cur;
                blk_next = (blk_cur < blk_end) ? P.blk[blk_cur].x - P.blk_j0 : INT32_MAX;
            }
            if (all_active || (j >= 0 && j <= T)) {
                uint2 X = make_uint2(0, 0);
                const int32_t *subcol = nullptr;
                if (SM == SCORE_PRMT) X = xtab[code];
                else subcol = subm + code;
                // START candidate value per cell (T5); NEG2 where START is out of scope
                const int sv_col = (start_any || (j == 0 && start_col0)) ? 0 : NEGK;
                // "G" = M + gap_open everywhere: the open penalty is paid once per cell
                // (it feeds both D of the next column and I of the next row) and the
                // substitution table holds s - gap_open, so diagG + s' == M_diag + s.
                int cm = INT32_MIN;   // column maximum over my rows (END_ANYWHERE)
                int capt = INT32_MIN; // G at lattice row Q (END_RESTRICTED)
                uint32_t w[WPL > 0 ? WPL : 1];
#pragma unroll
                for (int k = 0; k < WPL; ++k) w[k] = 0;
                // Phase A, rows BOTTOM-UP, all independent: D of this column and the
                // (match | D) candidate of M.  Going upwards lets every result overwrite
                // the register of the previous column's value it replaces (row r needs
                // the old M of row r-1 as its diagonal, which is still untouched), so the
                // loop-carried arrays are updated in place with no register copies.
                auto phase_a = [&](auto MASKED) {
                    constexpr bool masked = decltype(MASKED)::value;
#pragma unroll
                    for (int r = R - 1; r >= 0; --r) {
                        int sc;
                        if (SM == SCORE_PRMT) sc = prmt_sx(X.x, X.y, sel[r]);
                        else sc = subcol[sel[r]];
                        int diag = (r == 0) ? topMprev : Mp[r - 1];
                        if (masked && ((bmask >> r) & 1u)) diag = NEGK;   // T4 skipped at a blocked cell
                        if (!TB) {
                            Dp[r] = __viaddmax_s32(Dp[r], extD, Mp[r]);
                            Mp[r] = __viaddmax_s32(diag, sc, Dp[r]);   // max(match, D); START and I follow
                        } else {
                            // D: T0 extend (tag 4) first, T2 open (tag 0) replaces only if strictly greater
                            const int Dt = __viaddmax_s32(Dp[r], extD8t, Mp[r]);
                            const int dm = (Dt & ~7) | 1;
                            // START candidate (T5): 0 where START is in scope
                            int sv = LOCAL ? 0 : sv_col;
                            if (!LOCAL && r == 0 && first_row_lane && (start_row0 || j == 0)) sv = 0;
                            // M without I: T4 match|3, T5 start|2, T6 from D|1
                            const int mt = __vimax3_s32(diag + (sc * 8 + 3), sv + 2, dm);
                            Dp[r] = dm;
                            Mp[r] = mt;
                            w[r / 8] |= (uint32_t)((mt & 3) | (Dt & 4)) << (4 * (r % 8));
                        }
                    }
                };
                if (BLK && bmask) phase_a(std::true_type{});
                else phase_a(std::false_type{});
                // Phase B, rows TOP-DOWN: the vertical chain I -> M -> G.
                int upM = topM, upI = topI;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    // START candidate (T5): identically 0 for a local model
                    int sv = LOCAL ? 0 : sv_col;
                    if (!LOCAL && r == 0) {
                        // lattice row 0: START scope QUERY/EDGE allows it; corner always
                        if (first_row_lane && (start_row0 || j == 0)) sv = 0;
                    }
                    int Iv, Mv;
                    if (!TB) {
                        Iv = __viaddmax_s32(upI, extI, upM);
                        if (LOCAL) Mv = __vimax_s32_relu(Mp[r], Iv);  // full-rate VIMNMX.RELU
                        else Mv = __vimax3_s32(Mp[r], sv, Iv);
                    } else {
                        // I: T1 extend (tag 4) first, T3 open (tag 0)
                        const int It = __viaddmax_s32(upI, extI8t, upM);
                        Iv = It & ~7;
                        // M: the tagged max of phase A against T7 from I|0 (last in order)
                        const int mt = Mp[r];
                        const int m2 = max(mt, Iv);
                        Mv = m2 & ~7;
                        w[r / 8] ^= (uint32_t)(((mt ^ m2) & 3) | ((It & 4) << 1)) << (4 * (r % 8));
                    }
                    const int Gv = TB ? Mv + open8 : add_open(Mv, one, open);
                    Mp[r] = Gv;
                    upM = Gv;
                    upI = Iv;
                    if (LOCAL) cm = max(cm, Gv);
                    else if (r == rQ) capt = Gv;
                }
                botM = upM;
                botI = upI;
                topMprev = topM;
                if (TB) {
                    if (WPL == 4) *reinterpret_cast<uint4 *>(tbp) = make_uint4(w[0], w[1], w[2], w[3]);
                    else if (WPL == 2) *reinterpret_cast<uint2 *>(tbp) = make_uint2(w[0], w[1]);
                    else tbp[0] = w[0];
                }
                if (write_top) {
                    top_out[j] = make_int2(botM, botI);
                    // published in groups of columns: the fence waits for the hand-off stores and costs a good
                    // part of a step; a consumer trails by the lane skew + one group (short targets: small groups)
                    if (W > 1 && ((j & pub_mask) == pub_mask || j == T)) {
                        __threadfence_block();  // the rows are written before the counter moves
                        vprog[warp] = (long long)sweep * (T + 1) + j + 1;
                    }
                }
                // ---- END bookkeeping ----
                if (LOCAL) {
                    pend_cm = cm;
                    pend_j = j;
                } else if (j == T && end_colT) {
                    // whole last column is an END edge: rows in increasing order
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = row0 + r;
                        if (i <= Q && (Mp[r] > best || (Mp[r] == best && j < best_j))) {
                            best = Mp[r]; best_j = j; best_i = i;
                        }
                    }
                } else if (rQ >= 0 && (end_rowQ || (corner_only && j == T))) {
                    if (capt > best || (capt == best && j < best_j)) {
                        best = capt; best_j = j; best_i = Q;
                    }
                }
            }
            if (TB) tbp += 32 * WPL;
            // hand the strip's bottom row and the column code to the next lane
            const int nM = __shfl_up_sync(0xffffffffu, botM, 1);
            const int nI = __shfl_up_sync(0xffffffffu, botI, 1);
            const int nC = __shfl_up_sync(0xffffffffu, code, 1);
            if (lane > 0) {
                topM = nM;
                topI = nI;
                in_code = nC;
            }
        };

        // fill (lanes switch on one by one), steady state, drain
        const int fill_end = min(31, nsteps);
        const int steady_end = max(fill_end, min(T + 1, nsteps));
        int s = 0;
        for (; s < fill_end; ++s) step(s, std::false_type{});
        for (; s < steady_end; ++s) step(s, std::true_type{});
        for (; s < nsteps; ++s) step(s, std::false_type{});
        if (LOCAL) settle_pending();
        __syncwarp();
    }
    // lexicographic warp reduction: max score, then min j, then min i
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
        const bool take = (ob > best) || (ob == best && (oj < best_j || (oj == best_j && oi < best_i)));
        if (take) { best = ob; best_j = oj; best_i = oi; }
    }
    if (W > 1) {   // combine the warps' sweeps (idle warps carry INT32_MIN)
        if (lane == 0) { red[warp][0] = best; red[warp][1] = best_j; red[warp][2] = best_i; }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 1; w < W; ++w) {
                const int ob = red[w][0], oj = red[w][1], oi = red[w][2];
                if ((ob > best) || (ob == best && (oj < best_j || (oj == best_j && oi < best_i)))) {
                    best = ob; best_j = oj; best_i = oi;
                }
            }
    }
    if (threadIdx.x == 0) {
        AffOut o;
        o.best = (best == INT32_MIN) ? best : (TB ? best / 8 : best) - open;  // tracked as G = M + open (TB: x 8)
        o.end_i = best_i;
        o.end_j = best_j;
        o.flags = (best == INT32_MIN) ? 1 : 0;
        outs[P.out_index] = o;
    }
}

}  // namespace c4b
