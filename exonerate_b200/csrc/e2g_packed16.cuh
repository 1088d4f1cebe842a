// e2g_packed16.cuh -- est2genome lattice fill with BOTH STRANDS of the model in one
// register: forward-strand states in the low 16 bits, reverse-strand states in the
// high 16 bits (DPX S16x2), one warp per lattice.
//
// Same closed model, candidate order and END rule as e2g_fill_kernel
// (e2g_systolic.cuh; reference: generated optimal:est2genome DP functions,
// src/c4/viterbi.c:1638-1727 over src/model/est2genome.c:57-93 +
// src/model/intron.c:588-697).  The forward and the reverse half of the model are
// the same automaton with different splice-site columns, so every max-plus
// instruction advances both.  Bit-identical to the int32 kernel whenever no
// halfword add can wrap: all reachable values lie in
//     [2 gap_open + intron_open - 300,  max_sub * min(Q, T)],
// and the host sends a batch here only when that fits 15 bits (e2g16_eligible).
//
// What changed against the int32 kernel, and why it is ~8x fewer instructions:
//  * two strands per instruction (S16x2);
//  * the two-column delay lines (intron open / close advance the target by 2) are
//    PING-PONG register arrays indexed by the parity of the step: the value of
//    column j overwrites column j-2 in place, no register shifting;
//  * the intron-start shadow (viterbi.c:413-422; read only by the length test of
//    intron.c:138-161) is carried as a saturating 16-bit AGE = j - start: "open"
//    sets 2, "loop" adds 1, the close candidate of column j is valid iff
//    age(j-2) + 2 >= min_intron (the host guarantees max_intron >= T + 2, so the
//    upper test can never fire; saturation keeps the lower one exact).  The age is
//    stored relative to that threshold, so "invalid" is its sign bit;
//  * which candidate won comes for free from VIMNMX.S16x2's predicate outputs
//    (__vibmax_s16x2), stored as raw "earlier candidate held" bits.
// Mapping: lane l owns 16 consecutive rows, lanes are skewed by one column, strips
// of 512 rows are swept one after the other with the hand-off row {G, I} in L2 --
// exactly the scheme of affine_systolic.cuh.
//
// Traceback record, one halfword per cell: per strand 7 bits (forward at bit 0,
// reverse at bit 7) + bit 14 = "END prefers the forward strand":
//   bit0 close >= match   bit1 M held against I   bit2 .. against D   bit3 .. against START
//   bit4 N: open held against loop   bit5 I: open held against extend   bit6 D: likewise
#pragma once
#include "affine_packed16.cuh"
#include "e2g_systolic.cuh"

namespace c4b {

constexpr int kE2pR = 16;  // rows per lane -> 512 rows per sweep

struct E2pPair {
    const uint8_t *q;    // query classes (PRMT) per position
    const uint8_t *t;    // target column codes per position
    const uint32_t *sp;  // per target position: int8 x4 {ss5_fwd, ss3_fwd, ss5_rev, ss3_rev}
    int32_t Q, T;
    uint16_t *tb;        // single-pass traceback: [sweep][step][lane][16] halfwords, or null
    uint2 *top;          // sweep hand-off rows {G, I}: [sweep boundary][T+1]; null when one sweep
    uint32_t *ck;        // column checkpoints [window-1][sweep][lane][row][7], or null
    int64_t out_index;
};

// Windowed traceback (find_path on long targets): pass 1 is the score-only fill, which
// also saves the complete column state {G, N, age (two columns each), D} of every row at
// the last column of each window of win_cols columns.  The traceback then refills ONLY
// the windows the path crosses, from the checkpoint to their left, with records -- and
// an intron is crossed in one jump, because the checkpoint holds its age (= length so
// far).  Record memory is 2 MB per lattice instead of 2 B per cell, so every lattice of
// a batch is resident at once and the refilled area is a few per cent of the lattice.
// Window width (E2gModel::win_cols, a power of two, per batch): the refill of a round covers at most one
// window under each cursor, so narrow windows refill fewer cells per exon crossed; the checkpoints (one
// register state per row per window) are what limits how narrow -- 512 columns when they fit a quarter
// of the record budget, else 1024 (e2g_batch_create; measured 256 / 512 / 1024: 585 / 595 / 573 GCUPS).
constexpr int kE2pWinMax = 1024;
constexpr int kE2pCkWords = 7;

struct E2pWalk {       // traceback cursor of one lattice between rounds
    int32_t i, j, state, x;      // state: 0 M, 1 I, 2 D, 3 N; x: 0 forward, 1 reverse
    int32_t n_runs, last_t, status, done;
};

enum { E2P_SCORE = 0, E2P_FULL_TB = 1, E2P_SCORE_CK = 2, E2P_WINDOW_TB = 3 };

// MODE: E2P_SCORE (END cell only), E2P_FULL_TB (records for the whole lattice),
// E2P_SCORE_CK (END cell + column checkpoints), E2P_WINDOW_TB (records for one window
// of one lattice, started from a checkpoint; active = list of lattices, walk = cursors).
// blockDim.x = 32 W (W <= kE2pMaxWarps): the W warps of a CTA take the sweeps (strips of 32 RR rows)
// of ONE lattice round-robin and run them as a pipeline -- sweep k+1 follows sweep k at >= 32 columns,
// reading the hand-off row the moment it is published (monotone counter in shared memory, published in
// groups of columns; the scheme of affine_fill_kernel) -- so a 1 kbp cDNA occupies several schedulers
// instead of one: small batches (a shard of the 1k-pair batch on 8 GPUs is 125 lattices) leave most of
// the GPU idle with one warp per lattice.
constexpr int kE2pMaxWarps = 8;

// PIPE = false is the one-warp kernel (W folds to 1 at compile time): measured on the B200, the pipelined
// forms win only while the batch leaves warp slots empty (profiles/r02_e2g_small.md; 16 rows on two warps,
// the first pipelined shape, lost from 500 lattices up and is no longer chosen by the host).
//
// RR = rows per lane.  16 (512-row sweeps) is the loaded-GPU shape.  8 / 4 (256- / 128-row sweeps) exist for
// SMALL batches: a 1 kbp cDNA becomes four / eight sweeps on as many pipelined warps, and -- because the sweeps
// of a window refill read the hand-off rows pass 1 left in L2 -- the window refills of a lattice run
// on independent warps as well.  All record / checkpoint layouts are [..][lane][RR]: one batch uses
// one RR throughout (E2gBatch::rows16).
template <int MODE, bool PIPE = false, int RR = kE2pR>
__global__ void __launch_bounds__(PIPE ? 32 * (RR == 4 ? 8 : 4) : 32)
e2g_fill16_kernel(const E2pPair *__restrict__ pairs, E2gOut *__restrict__ outs, const E2gModel mdl,
                  const uint2 *__restrict__ score_table, const int32_t *__restrict__ active,
                  const E2pWalk *__restrict__ walk, uint16_t *__restrict__ winbuf, size_t win_stride) {
    constexpr int R = RR;
    static_assert(R == 4 || R == 8 || R == 16, "records are stored as uint4 groups of 8 rows (uint2 for 4)");
    constexpr bool TB = (MODE == E2P_FULL_TB || MODE == E2P_WINDOW_TB);
    constexpr bool WIN = (MODE == E2P_WINDOW_TB);
    constexpr bool CK = (MODE == E2P_SCORE_CK);
    __shared__ uint2 xtab[25];
    __shared__ volatile long long vprog[kE2pMaxWarps];
    __shared__ int red[kE2pMaxWarps][6];
    const int lane = threadIdx.x & 31, warp = PIPE ? (int)(threadIdx.x >> 5) : 0, W = PIPE ? (int)(blockDim.x >> 5) : 1;
    const int pidx = WIN ? active[blockIdx.x] : (int)blockIdx.x;
    const E2pPair P = pairs[pidx];
    const int Q = P.Q, T = P.T;
    if (threadIdx.x < 25) xtab[threadIdx.x] = score_table[threadIdx.x];
    if (PIPE) {
        if (threadIdx.x < kE2pMaxWarps) vprog[threadIdx.x] = 0;
        __syncthreads();
    } else {
        __syncwarp();
    }

    const uint32_t open2 = pack16(mdl.open), ext2 = pack16(mdl.ext);
    const uint32_t preK = pack16(mdl.intron_open - mdl.open);  // N opens from G = M + open
    // ages are stored relative to the validity threshold: a = (j - start) - (min_intron - 2)
    const int age_thr = max(0, mdl.min_intron - 2);
    const uint32_t openA = pack16(2 - age_thr);
    const int rows_per_sweep = 32 * R;
    const int all_sweeps = (Q + 1 + rows_per_sweep - 1) / rows_per_sweep;
    const int nsteps = T + 1 + 31;
    // the fence of a publish waits for the hand-off stores: groups of 32 columns (8 on short targets,
    // where the pipeline's run-in counts)
    const int pub_mask = (T >= 4096) ? 31 : 7;
    // column range [c0, c1] and sweeps of this launch
    int c0 = 0, c1 = T, nsweeps = all_sweeps;
    if (WIN) {
        const E2pWalk W = walk[pidx];
        c0 = W.j & ~(mdl.win_cols - 1);
        c1 = W.j;                                   // the path never moves right or down
        nsweeps = min(all_sweeps, W.i / rows_per_sweep + 1);
    }
    const uint32_t *ck_in = (WIN && c0 > 0) ? P.ck + (size_t)(c0 / mdl.win_cols - 1) * all_sweeps * 32 * R * kE2pCkWords
                                            : nullptr;

    // first strict maximum of END per strand, tracked on G = M + open
    int bestF = INT32_MIN, bjF = 0, biF = 0, bestR = INT32_MIN, bjR = 0, biR = 0;
    uint32_t best2 = kMin16x2;

    for (int sweep = warp; sweep < nsweeps; sweep += W) {
        const int row0 = sweep * rows_per_sweep + lane * R;
        const bool first_row_lane = (sweep == 0 && lane == 0);
        const bool later_sweep = (sweep > 0);
        // END bookkeeping is per warp: an earlier sweep OF THIS WARP may hold a tie at a larger column
        const bool later_mine = (sweep != warp);
        // the sweep above may still be running on another warp
        // (a window refill never waits: its hand-off rows are the ones pass 1 kept)
        const bool piped = PIPE && !WIN && later_sweep && W > 1;
        const int wp = (sweep - 1) % W;
        const long long in_base = (long long)(sweep - 1) * (T + 1);
        long long avail = 0;
        auto wait_column = [&](int col) {   // until column `col` of the sweep above is published
            const long long need = in_base + col + 1;
            if (avail < need) {
                while ((avail = vprog[wp]) < need) __nanosleep(40);
                __threadfence_block();
            }
        };
        const int nvalid = min(R, max(0, Q - row0 + 1));  // rows r < nvalid are lattice rows <= Q
        uint32_t sel[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = row0 + r;
            const uint32_t c = (i >= 1 && i <= Q) ? (uint32_t)P.q[i - 1] : (uint32_t)kPadClass;
            const uint32_t s = c | ((c | 8u) << 4);  // byte c, sign-extended
            sel[r] = s | (s << 8);                   // the same score in both halves
        }
        // ping-pong by step parity: [p] holds the column two steps back and is overwritten
        uint32_t G[2][R], N[2][R], A[2][R], Dp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            G[0][r] = G[1][r] = N[0][r] = N[1][r] = Dp[r] = kNeg16x2;
            A[0][r] = A[1][r] = 0u;
        }
        if (WIN && ck_in) {
            // columns c0-1 ("new") and c0-2 ("old") of my rows; my first step is c0 + lane,
            // and the array indexed by a step's parity must hold the column two steps back
            const uint32_t *c = ck_in + ((size_t)(sweep * 32 + lane) * R) * kE2pCkWords;
            const int po = (c0 + lane) & 1;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t gn = c[r * kE2pCkWords + 0], go = c[r * kE2pCkWords + 1];
                const uint32_t nn = c[r * kE2pCkWords + 2], no = c[r * kE2pCkWords + 3];
                const uint32_t an = c[r * kE2pCkWords + 4], ao = c[r * kE2pCkWords + 5];
                Dp[r] = c[r * kE2pCkWords + 6];
                G[0][r] = po ? gn : go; G[1][r] = po ? go : gn;
                N[0][r] = po ? nn : no; N[1][r] = po ? no : nn;
                A[0][r] = po ? an : ao; A[1][r] = po ? ao : an;
            }
        }
        const uint2 *top_in = later_sweep ? P.top + (size_t)(sweep - 1) * (T + 1) : nullptr;  // written by sweep-1
        uint2 *top_out = (!WIN && sweep + 1 < all_sweeps) ? P.top + (size_t)sweep * (T + 1) : nullptr;
        const bool write_top = (top_out != nullptr) && (lane == 31);
        uint32_t topG = kNeg16x2, topI = kNeg16x2, topGprev = kNeg16x2;
        // lane 0's inputs for its first column c0: symbol, splice word of column c0-2, row above
        int in_code = kTargetNone, code0 = (c0 >= 1) ? (int)P.t[c0 - 1] : kTargetNone;
        uint32_t in_sp = 0u, sp0 = (c0 >= 2) ? P.sp[c0 - 2] : 0u;
        uint2 top0v = make_uint2(kNeg16x2, kNeg16x2);
        // hand-off columns are fetched PF steps ahead of their use (tq = columns s+1 .. s+PF-1).  A step
        // of 8 rows is shorter than an L2 round trip, but PF = 4 measured 4 % SLOWER than 1 on the B200
        // (125 lattices 238 -> 228 GCUPS: the queue moves cost more than the latency they hide; ncu
        // puts the pipelined shape's loss in exposed branch / fixed-latency stalls of a warp that is
        // alone on its scheduler, profiles/r02_e2g_small.md), so the depth stays 1
        constexpr int PF = 1;
        uint2 tq[PF > 1 ? PF - 1 : 1];
        // diagonal input of my first row at column c0: G of the row above at column c0-1
        if (WIN && ck_in && lane > 0)
            topGprev = ck_in[((size_t)(sweep * 32 + lane - 1) * R + (R - 1)) * kE2pCkWords + 0];
        if (later_sweep) {
            if (piped) wait_column(min(c0 + PF - 1, T));
            top0v = PIPE ? __ldcg(top_in + c0) : top_in[c0];
#pragma unroll
            for (int k = 0; k < PF - 1; ++k)
                tq[k] = (c0 + 1 + k <= T) ? (PIPE ? __ldcg(top_in + c0 + 1 + k) : top_in[c0 + 1 + k])
                                          : make_uint2(kNeg16x2, kNeg16x2);
            if (c0 >= 1 && lane == 0) topGprev = top_in[c0 - 1].x;
        }
        uint16_t *tbp = nullptr;   // this lane's R halfwords of the current step
        if (MODE == E2P_FULL_TB) tbp = P.tb + (((size_t)sweep * nsteps) * 32 + lane) * R;
        if (WIN) tbp = winbuf + (size_t)blockIdx.x * win_stride + (((size_t)sweep * (mdl.win_cols + 31)) * 32 + lane) * R;

        auto step = [&](const int s, auto PAR) {
            constexpr int p = decltype(PAR)::value, o = p ^ 1;
            const int j = s - lane;
            const int code = (lane == 0) ? code0 : in_code;
            const uint32_t spw = (lane == 0) ? sp0 : in_sp;  // splice word of column j-2
            if (later_sweep && lane == 0) { topG = top0v.x; topI = top0v.y; }
            if (s + 1 <= T) {
                if constexpr (R <= 8) {
                    // the byte as a 32-bit load result: nvcc otherwise masks it (LOP3 & 0xff) right behind the
                    // load, and a warp that is alone on its scheduler then sits out the whole load latency in
                    // every step (ncu: 10 % of the small-batch kernel's samples on that one instruction)
                    unsigned v;
                    asm("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(P.t + s));
                    code0 = (int)v;
                } else {
                    code0 = (int)P.t[s];
                }
                sp0 = (s >= 1) ? P.sp[s - 1] : 0u;   // source column (s+1)-2
                if (later_sweep) {
                    uint2 nv = make_uint2(kNeg16x2, kNeg16x2);
                    if (s + PF <= T) {
                        if (piped) wait_column(s + PF);
                        nv = PIPE ? __ldcg(top_in + s + PF) : top_in[s + PF];
                    }
                    if constexpr (PF > 1) {
                        top0v = tq[0];
#pragma unroll
                        for (int k = 0; k + 1 < PF - 1; ++k) tq[k] = tq[k + 1];
                        tq[PF - 2] = nv;
                    } else {
                        top0v = nv;
                    }
                }
            } else {
                code0 = kTargetNone;
                sp0 = 0u;
            }
            uint32_t botG = kNeg16x2, botI = kNeg16x2;
            if (j >= c0 && j <= c1) {
                const uint2 X = xtab[code];
                // forward opens at a 5' site and closes at a 3' site, reverse 3' then 5'
                uint32_t pre2, post2;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(pre2) : "r"(spw), "r"(0u), "r"(0xB380u));
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(post2) : "r"(spw), "r"(0u), "r"(0xA291u));
                pre2 = __vadd2(pre2, preK);
                uint32_t rec[TB ? R : 1];
                // ---- phase A: everything that depends on previous columns only ----------
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    uint32_t sc;
                    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sc) : "r"(X.x), "r"(X.y), "r"(sel[r]));
                    // N: open (T3/T0) first, loop (T4/T1) replaces only if strictly greater.  The
                    // per-strand choice is a halfword MASK (sign of open - loop, replicated by one
                    // PRMT), so the age follows with one LOP3 and no predicates
                    const uint32_t nv = __vadd2(G[p][r], pre2);
                    const uint32_t lm = sign_mask16x2(__vsub2(nv, N[o][r]));  // 0xFFFF: loop replaces open
                    const uint32_t Nn = __vmaxs2(nv, N[o][r]);
                    uint32_t An = __viaddmin_s16x2(A[o][r], 0x00010001u, 0x7FFE7FFEu);   // age + 1, saturating
                    An = (An & lm) | (openA & ~lm);
                    // close candidate (T5/T2) from column j-2, valid iff its intron is long enough:
                    // the age is kept relative to the threshold, invalid <=> negative
                    const uint32_t vm = sign_mask16x2(A[p][r]);
                    uint32_t c5 = __vadd2(N[p][r], post2);
                    c5 = (c5 & ~vm) | (kNeg16x2 & vm);
                    const bool ol = !(lm & 1u), oh = !(lm >> 31);
                    const uint32_t diag = (r == 0) ? topGprev : G[o][r - 1];
                    uint32_t xc, Dn;
                    if (!TB) {
                        xc = __viaddmax_s16x2(diag, sc, c5);               // max(close, match)
                        Dn = __viaddmax_s16x2(Dp[r], ext2, G[o][r]);       // max(open, extend)
                    } else {
                        bool ch, cl, dh, dl;
                        xc = __vibmax_s16x2(c5, __vadd2(diag, sc), &ch, &cl);       // close first (T5, T11)
                        Dn = __vibmax_s16x2(G[o][r], __vadd2(Dp[r], ext2), &dh, &dl);  // open first (T13, T15)
                        rec[r] = (cl ? 0x0001u : 0u) | (ch ? 0x0080u : 0u) | (ol ? 0x0010u : 0u) | (oh ? 0x0800u : 0u) |
                                 (dl ? 0x0040u : 0u) | (dh ? 0x2000u : 0u);
                    }
                    N[p][r] = Nn;
                    A[p][r] = An;
                    Dp[r] = Dn;
                    G[p][r] = xc;  // until phase B turns it into G of this column
                }
                // ---- phase B: the vertical chain I -> M -> G, top-down ---------------------
                uint32_t upG = topG, upI = topI, cm = kMin16x2;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    uint32_t Iv, Mv;
                    if (!TB) {
                        Iv = __viaddmax_s16x2(upI, ext2, upG);
                        Mv = __vimax3_s16x2_relu(G[p][r], Iv, Dp[r]);
                    } else {
                        bool ih, il, ah, al, bh, bl, sh, sl;
                        Iv = __vibmax_s16x2(upG, __vadd2(upI, ext2), &ih, &il);   // open first (T12, T14)
                        Mv = __vibmax_s16x2(G[p][r], Iv, &ah, &al);               // then I (T19/T16)
                        Mv = __vibmax_s16x2(Mv, Dp[r], &bh, &bl);                 // then D (T20/T17)
                        Mv = __vibmax_s16x2(Mv, 0u, &sh, &sl);                    // then START (T21/T18)
                        // END: reverse strand first (T22), forward (T23) only if strictly greater
                        const bool endf = lo16(Mv) > hi16(Mv);
                        rec[r] |= (al ? 0x0002u : 0u) | (ah ? 0x0100u : 0u) | (bl ? 0x0004u : 0u) | (bh ? 0x0200u : 0u) |
                                  (sl ? 0x0008u : 0u) | (sh ? 0x0400u : 0u) | (il ? 0x0020u : 0u) | (ih ? 0x1000u : 0u) |
                                  (endf ? 0x4000u : 0u);
                    }
                    const uint32_t Gv = __vadd2(Mv, open2);
                    G[p][r] = Gv;
                    upG = Gv;
                    upI = Iv;
                    if (r & 1) cm = __vimax3_s16x2(cm, Gv, G[p][r - 1]);
                }
                botG = upG;
                botI = upI;
                topGprev = topG;
                if (TB) {
                    if constexpr (R >= 8) {
#pragma unroll
                        for (int g = 0; g < R / 8; ++g)
                            reinterpret_cast<uint4 *>(tbp)[g] =
                                make_uint4(rec[8 * g] | (rec[8 * g + 1] << 16), rec[8 * g + 2] | (rec[8 * g + 3] << 16),
                                           rec[8 * g + 4] | (rec[8 * g + 5] << 16), rec[8 * g + 6] | (rec[8 * g + 7] << 16));
                    } else {
                        *reinterpret_cast<uint2 *>(tbp) = make_uint2(rec[0] | (rec[1] << 16), rec[2] | (rec[3] << 16));
                    }
                }
                if (write_top) {
                    top_out[j] = make_uint2(botG, botI);
                    if (PIPE && W > 1 && ((j & pub_mask) == pub_mask || j == T)) {
                        __threadfence_block();   // the rows are written before the counter moves
                        vprog[warp] = (long long)sweep * (T + 1) + j + 1;
                    }
                }
                if (CK && ((j + 1) & (mdl.win_cols - 1)) == 0 && j < T) {
                    // last column of a window: the state a later window refill starts from
                    uint32_t *c = P.ck + ((size_t)(((j + 1) / mdl.win_cols - 1) * all_sweeps + sweep) * 32 + lane) * R * kE2pCkWords;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        c[r * kE2pCkWords + 0] = G[p][r]; c[r * kE2pCkWords + 1] = G[o][r];
                        c[r * kE2pCkWords + 2] = N[p][r]; c[r * kE2pCkWords + 3] = N[o][r];
                        c[r * kE2pCkWords + 4] = A[p][r]; c[r * kE2pCkWords + 5] = A[o][r];
                        c[r * kE2pCkWords + 6] = Dp[r];
                    }
                }
                // ---- END bookkeeping: cm covers padding rows too, so it only TRIGGERS the exact
                // search (first sweep: strictly greater; later sweeps: a tie at a smaller j wins)
                bool trig = false;
                if (!WIN) {
                    bool gh, gl;
                    if (later_mine) {
                        (void)__vibmax_s16x2(cm, best2, &gh, &gl);   // cm >= best
                        trig = gh || gl;
                    } else {
                        (void)__vibmax_s16x2(best2, cm, &gh, &gl);   // best >= cm
                        trig = !(gh && gl);
                    }
                }
                if (trig) {
                    int vF = INT32_MIN, iF = 0, vR = INT32_MIN, iR = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (r < nvalid) {
                            const int f = lo16(G[p][r]), rv = hi16(G[p][r]);
                            if (f > vF) { vF = f; iF = r; }
                            if (rv > vR) { vR = rv; iR = r; }
                        }
                    if (vF > bestF || (vF == bestF && j < bjF)) { bestF = vF; bjF = j; biF = row0 + iF; }
                    if (vR > bestR || (vR == bestR && j < bjR)) { bestR = vR; bjR = j; biR = row0 + iR; }
                    best2 = ((uint32_t)max(bestF, -32768) & 0xFFFFu) | ((uint32_t)max(bestR, -32768) << 16);
                }
            }
            if (first_row_lane) { topG = kNeg16x2; topI = kNeg16x2; }
            if (TB) tbp += 32 * R;
            const uint32_t nG = __shfl_up_sync(0xffffffffu, botG, 1);
            const uint32_t nI = __shfl_up_sync(0xffffffffu, botI, 1);
            const int nC = __shfl_up_sync(0xffffffffu, code, 1);
            const uint32_t nS = __shfl_up_sync(0xffffffffu, spw, 1);
            if (lane > 0) {
                topG = nG;
                topI = nI;
                in_code = nC;
                in_sp = nS;
            }
        };

        // steps c0 .. c1+31 (lane l works on column step - l); parity = step & 1
        int s = c0;
        const int s_end = c1 + 32;
        if (s & 1) { step(s, std::integral_constant<int, 1>{}); ++s; }
        for (; s + 1 < s_end; s += 2) {
            step(s, std::integral_constant<int, 0>{});
            step(s + 1, std::integral_constant<int, 1>{});
        }
        if (s < s_end) step(s, std::integral_constant<int, 0>{});
        __syncwarp();
    }
    if constexpr (!WIN) {

    // per strand: lexicographic warp reduction (max score, min j, min i)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bestF, off), oj = __shfl_xor_sync(0xffffffffu, bjF, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biF, off);
            if (ob > bestF || (ob == bestF && (oj < bjF || (oj == bjF && oi < biF)))) { bestF = ob; bjF = oj; biF = oi; }
        }
        {
            const int ob = __shfl_xor_sync(0xffffffffu, bestR, off), oj = __shfl_xor_sync(0xffffffffu, bjR, off);
            const int oi = __shfl_xor_sync(0xffffffffu, biR, off);
            if (ob > bestR || (ob == bestR && (oj < bjR || (oj == bjR && oi < biR)))) { bestR = ob; bjR = oj; biR = oi; }
        }
    }
    if (PIPE && W > 1) {   // combine the warps' sweeps (idle warps carry INT32_MIN)
        if (lane == 0) {
            red[warp][0] = bestF; red[warp][1] = bjF; red[warp][2] = biF;
            red[warp][3] = bestR; red[warp][4] = bjR; red[warp][5] = biR;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 1; w < W; ++w) {
                int ob = red[w][0], oj = red[w][1], oi = red[w][2];
                if (ob > bestF || (ob == bestF && (oj < bjF || (oj == bjF && oi < biF)))) { bestF = ob; bjF = oj; biF = oi; }
                ob = red[w][3]; oj = red[w][4]; oi = red[w][5];
                if (ob > bestR || (ob == bestR && (oj < bjR || (oj == bjR && oi < biR)))) { bestR = ob; bjR = oj; biR = oi; }
            }
    }
    if (threadIdx.x == 0) {
        // the first cell (target outer, query inner) that reaches the overall maximum; in that
        // cell the reverse strand is tried first and the forward one must be strictly greater
        bool fwd;
        if (bestF != bestR) fwd = bestF > bestR;
        else if (bjF != bjR) fwd = bjF < bjR;
        else if (biF != biR) fwd = biF < biR;
        else fwd = false;
        E2gOut o;
        o.best = (fwd ? bestF : bestR) - mdl.open;
        o.end_i = fwd ? biF : biR;
        o.end_j = fwd ? bjF : bjR;
        o.end_forward = fwd ? 1 : 0;
        outs[P.out_index] = o;
    }
    }
}

// Viterbi_Data_create_Alignment (viterbi.c:342-392) over the 15-bit records.
__global__ void e2g16_traceback_kernel(const E2pPair *__restrict__ pairs, const E2gOut *__restrict__ outs,
                                       const E2gJob *__restrict__ jobs, int n, const E2gModel mdl, int threshold,
                                       c4b_result *__restrict__ results, int32_t *__restrict__ ops, int R) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const E2gJob J = jobs[g];
    const E2pPair P = pairs[J.pair];
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
        const int w = ci / (32 * R), ln = (ci / R) & 31, r = ci % R;
        return P.tb[(((size_t)w * nsteps + (cj + ln)) * 32 + ln) * R + r];
    };
    if (o.best < threshold) {
        res.status = 1;
    } else {
        const int x = o.end_forward ? 0 : 1;  // 0 = forward fields (bits 0..6), 1 = reverse (bits 7..13)
        int state = 0;                         // 0 M, 1 I, 2 D, 3 N
        emit(mdl.tM2E[x]);
        for (;;) {
            const uint32_t f = (record(i, j) >> (7 * x)) & 127u;
            if (state == 0) {
                if (!(f & 8u)) { emit(mdl.tS2M[x]); break; }
                else if (!(f & 4u)) { emit(mdl.tD2M[x]); state = 2; }
                else if (!(f & 2u)) { emit(mdl.tI2M[x]); state = 1; }
                else if (f & 1u) { emit(mdl.tNclose[x]); j -= 2; state = 3; }
                else { emit(mdl.tMatch[x]); --i; --j; }
            } else if (state == 1) {
                if (f & 32u) { emit(mdl.tIopen[x]); state = 0; } else emit(mdl.tIext[x]);
                --i;
            } else if (state == 2) {
                if (f & 64u) { emit(mdl.tDopen[x]); state = 0; } else emit(mdl.tDext[x]);
                --j;
            } else {
                if (f & 16u) { emit(mdl.tNopen[x]); j -= 2; state = 0; }
                else { emit(mdl.tNloop[x]); --j; }
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

// ---- windowed traceback: cursors, window records, intron jumps -----------------------
__global__ void e2g16_walk_init_kernel(const E2pPair *__restrict__ pairs, const E2gOut *__restrict__ outs,
                                       const E2gJob *__restrict__ jobs, int n, const E2gModel mdl, int threshold,
                                       E2pWalk *__restrict__ walk, int32_t *__restrict__ ops) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const E2gJob J = jobs[g];
    const E2gOut o = outs[pairs[J.pair].out_index];
    E2pWalk W;
    W.i = o.end_i; W.j = o.end_j; W.state = 0; W.x = o.end_forward ? 0 : 1;
    W.n_runs = 0; W.last_t = -1; W.status = 0; W.done = 0;
    if (o.best < threshold) { W.status = 1; W.done = 1; }
    else if (J.ops_cap >= 1) {
        int32_t *out = ops + 2 * J.ops_off;
        out[0] = mdl.tM2E[W.x]; out[1] = 1; W.n_runs = 1; W.last_t = mdl.tM2E[W.x];
    } else { W.status = 4; W.done = 1; }
    walk[J.pair] = W;
}

// which lattices still have a traceback in flight (order = job order)
__global__ void e2g16_active_kernel(const E2pWalk *__restrict__ walk, int n, int32_t *__restrict__ active,
                                    int32_t *__restrict__ count) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    if (!walk[g].done) active[atomicAdd(count, 1)] = g;
}

// Viterbi_Data_create_Alignment (viterbi.c:342-392) inside the window that was just
// refilled; leaves the cursor at the first cell left of the window (or finishes).
__global__ void e2g16_walk_kernel(const E2pPair *__restrict__ pairs, const E2gOut *__restrict__ outs,
                                  const E2gJob *__restrict__ jobs, const int32_t *__restrict__ active, int n_active,
                                  const E2gModel mdl, E2pWalk *__restrict__ walk,
                                  const uint16_t *__restrict__ winbuf, size_t win_stride,
                                  c4b_result *__restrict__ results, int32_t *__restrict__ ops, int R) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_active) return;
    const int pidx = active[slot];
    const E2gJob J = jobs[pidx];   // jobs are indexed like pairs (J.pair == pidx)
    const E2pPair P = pairs[pidx];
    E2pWalk W = walk[pidx];
    const int c0 = W.j & ~(mdl.win_cols - 1);
    const int all_sweeps = (P.Q + 1 + 32 * R - 1) / (32 * R);
    const uint16_t *rec_base = winbuf + (size_t)slot * win_stride;
    int32_t *out = ops + 2 * J.ops_off;
    int i = W.i, j = W.j, state = W.state, n_runs = W.n_runs, last_t = W.last_t;
    const int x = W.x;
    bool overflow = false, finished = false;
    auto emit = [&](int t, int cnt) {
        if (cnt <= 0) return;
        if (t == last_t) out[2 * (n_runs - 1) + 1] += cnt;
        else if (n_runs < J.ops_cap) { out[2 * n_runs] = t; out[2 * n_runs + 1] = cnt; ++n_runs; last_t = t; }
        else overflow = true;
    };
    auto record = [&](int ci, int cj) -> uint32_t {
        const int w = ci / (32 * R), ln = (ci / R) & 31, r = ci % R;
        return rec_base[(((size_t)w * (mdl.win_cols + 31) + (cj - c0 + ln)) * 32 + ln) * R + r];
    };
    while (j >= c0) {
        const uint32_t f = (record(i, j) >> (7 * x)) & 127u;
        if (state == 0) {
            if (!(f & 8u)) { emit(mdl.tS2M[x], 1); finished = true; break; }
            else if (!(f & 4u)) { emit(mdl.tD2M[x], 1); state = 2; }
            else if (!(f & 2u)) { emit(mdl.tI2M[x], 1); state = 1; }
            else if (f & 1u) { emit(mdl.tNclose[x], 1); j -= 2; state = 3; }
            else { emit(mdl.tMatch[x], 1); --i; --j; }
        } else if (state == 1) {
            if (f & 32u) { emit(mdl.tIopen[x], 1); state = 0; } else emit(mdl.tIext[x], 1);
            --i;
        } else if (state == 2) {
            if (f & 64u) { emit(mdl.tDopen[x], 1); state = 0; } else emit(mdl.tDext[x], 1);
            --j;
        } else {
            if (f & 16u) { emit(mdl.tNopen[x], 1); j -= 2; state = 0; }
            else { emit(mdl.tNloop[x], 1); --j; }
        }
        if (i < 0 || j < 0 || overflow) break;
    }
    if (!finished && !overflow && i >= 0 && j >= 0 && state == 3 && c0 > 0) {
        // inside an intron at column j in {c0-1, c0-2}: the checkpoint left of this window
        // holds the intron's age there = j - (column the intron was opened from)
        const int w = i / (32 * R), ln = (i / R) & 31, r = i % R;
        const uint32_t *c = P.ck + (((size_t)(c0 / mdl.win_cols - 1) * all_sweeps + w) * 32 + ln) * R * kE2pCkWords +
                            (size_t)r * kE2pCkWords;
        const uint32_t a2 = (j == c0 - 1) ? c[4] : c[5];
        const int rel = (int)(short)((a2 >> (16 * x)) & 0xFFFFu);   // stored relative to the threshold
        const int age = rel + max(0, mdl.min_intron - 2);
        if (rel < 0x7FFE && age >= 2 && age <= j) {
            emit(mdl.tNloop[x], age - 2);
            emit(mdl.tNopen[x], 1);
            j -= age;
            state = 0;
        }
    }
    if (finished || overflow || i < 0 || j < 0) {
        c4b_result res;
        const E2gOut o = outs[P.out_index];
        res.score = o.best; res.reserved = 0; res.ops_offset = J.ops_off;
        res.status = finished && !overflow ? 0 : 4;
        res.query_end = J.q_origin + o.end_i;
        res.target_end = J.t_origin + o.end_j;
        if (res.status == 0)
            for (int a = 0, b = n_runs - 1; a < b; ++a, --b) {
                const int t0 = out[2 * a], l0 = out[2 * a + 1];
                out[2 * a] = out[2 * b]; out[2 * a + 1] = out[2 * b + 1];
                out[2 * b] = t0; out[2 * b + 1] = l0;
            }
        res.n_ops = (res.status == 0) ? n_runs : 0;
        res.query_start = J.q_origin + max(i, 0);
        res.target_start = J.t_origin + max(j, 0);
        results[J.result] = res;
        W.done = 1;
        W.status = res.status;
    }
    W.i = i; W.j = j; W.state = state; W.n_runs = n_runs; W.last_t = last_t;
    walk[pidx] = W;
}

// lattices whose score is below the threshold never enter the rounds
__global__ void e2g16_walk_rejected_kernel(const E2pPair *__restrict__ pairs, const E2gOut *__restrict__ outs,
                                           const E2gJob *__restrict__ jobs, int n, const E2pWalk *__restrict__ walk,
                                           c4b_result *__restrict__ results) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const E2gJob J = jobs[g];
    const E2pWalk W = walk[J.pair];
    if (!(W.done && W.status != 0 && W.n_runs == 0)) return;
    const E2gOut o = outs[pairs[J.pair].out_index];
    c4b_result res;
    res.score = o.best; res.status = W.status; res.reserved = 0; res.n_ops = 0; res.ops_offset = J.ops_off;
    res.query_end = J.q_origin + o.end_i; res.target_end = J.t_origin + o.end_j;
    res.query_start = J.q_origin + o.end_i; res.target_start = J.t_origin + o.end_j;
    results[J.result] = res;
}

__global__ void e2g16_score_results_kernel(const E2pPair *__restrict__ pairs, const E2gOut *__restrict__ outs,
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
