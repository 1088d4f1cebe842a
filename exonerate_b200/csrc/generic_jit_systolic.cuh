// generic_jit_systolic.cuh -- the table-driven lattice fill as a SYSTOLIC kernel,
// specialised to one closed C4 model at run time (NVRTC; never seen by nvcc).
//
// generic_jit_kernel.cuh gives every lattice row a thread and walks anti-diagonals
// with a CTA-wide barrier per diagonal, the lattice in a shared-memory ring: 635
// warp-instructions per warp-cell for protein2genome, a third of them ring
// addressing (profiles/r01d_generic_jit.md).  This file is the mapping the
// hand-written affine / est2genome kernels use (affine_systolic.cuh), generated for
// ANY closed model:
//   * a lane owns R consecutive lattice rows; lanes are skewed by one column (lane l
//     works on column step - l), so the warp is the anti-diagonal wavefront;
//   * the lattice lives in REGISTERS: per row, the states that are ever read at a
//     non-zero advance keep their last advance_target + 1 columns (V[row][...], shifted
//     once per step); nothing of the lattice touches shared memory or HBM;
//   * the rows above a lane's strip (as many as the model's largest advance_query) are
//     "virtual rows" of the same register file, fed by one warp shuffle per carried word
//     per step from the lane above; queries longer than 32 R rows are swept in strips
//     by the W warps of the CTA, pipelined, the strip hand-off row in L2 exactly as in
//     affine_fill_kernel (monotone counter in shared memory, no __syncthreads);
//   * shadow slots are carried only by the states that lie between a shadow's start and
//     the transitions that read it (kNeed, computed on the host from the closed model).
// Cell semantics are those of Viterbi_interpreted (src/c4/viterbi.c:655-837): per cell
// every state starts unset, transitions are tried in closed-model order, the first valid
// one assigns, later ones replace only if strictly greater (:766-775); END is the first
// strict maximum in (target outer, query inner) order (:778-791).  calc_score<> and the
// scope tests are the ones of generic_jit_kernel.cuh (included in front of this file).
//
// SubOpt blocked cells (JIT_SYS_BLK; src/c4/subopt.h:77-80, viterbi.c:701-704: a MATCH-labelled
// transition is skipped at a blocked DESTINATION cell): the host turns a lattice's list into {column,
// row mask} entries per lane strip, sorted by column (GenPair::blk / blk_off); a lane walks its strip's
// entries with one cursor and the mask of the current column gates the match transitions of its rows --
// the scheme of affine_fill_kernel's BLK variant.
// Not handled here (the host keeps generic_jit_kernel.cuh for them): START / END cell tables of
// BSDP's derived models.
//
// JIT_SYS_WIN -- a PATH record for lattices whose whole record would not fit device memory (the
// reference recurses through checkpoint rows for the same reason, src/c4/optimal.c:183-345,
// src/c4/viterbi.c:515-631; results do not depend on it):
//   1  whole lattice (score), and every lane saves its complete register lattice V after the last
//      column of each window of `wcols` columns (a column checkpoint: everything the columns to its
//      right depend on);
//   2  PATH over ONE window of columns [c0, c1] and the strips down to the traceback cursor,
//      started from the checkpoint to its left; records are laid out as a lattice of wcols columns.
// The host alternates window refills and a resumable walk (generic_window_walk_kernel) from the
// END cell leftwards: record memory is one window, not the lattice.
//
// Expected in front of this file, after generic_jit_kernel.cuh's own tables:
//   JIT_SYS_R (rows per lane), namespace c4bjit { AQ (largest advance_query), VW (words
//   per row), kNW[S] (words a state carries), kVD[S] (columns kept - 1; -1: never stored),
//   kVOff[S], kNeed[S * C4B_MAX_SHADOW_SLOTS], NSEND (words handed down per step),
//   kSendD[NSEND], kSendOff[NSEND], kTbCode[TN], kTbBits[S], kTbBitOff[S], TB_ROW_BITS,
//   TB_CHUNK (PATH record, see the host's SysLayout) }.

#ifndef JIT_SYS_WIN
#define JIT_SYS_WIN 0
#endif
#ifndef JIT_SYS_BLK
#define JIT_SYS_BLK 0
#endif

namespace c4bjit {

constexpr int kWinMode = JIT_SYS_WIN;
constexpr bool kBlk = JIT_SYS_BLK != 0;
static_assert(kWinMode == 0 || (kWinMode == 1 && JIT_MODE == GEN_SCORE) || (kWinMode == 2 && JIT_MODE == GEN_PATH),
              "checkpoints are written by the score pass and consumed by the PATH pass");
constexpr int SR = JIT_SYS_R;
constexpr int NROWS = AQ + SR;        // virtual rows above the strip first
constexpr int kSysMaxWarps = 8;
// word index of shadow slot l / of the packed START cell inside a state's carried words
__device__ constexpr int word_of_slot(int s, int l) {
    int w = 1;
    for (int k = 0; k < l; ++k) w += kNeed[s * C4B_MAX_SHADOW_SLOTS + k] ? 1 : 0;
    return kNeed[s * C4B_MAX_SHADOW_SLOTS + l] ? w : -1;
}
__device__ constexpr int word_of_start(int s) {
    int w = 1;
    for (int k = 0; k < NSH; ++k) w += kNeed[s * C4B_MAX_SHADOW_SLOTS + k] ? 1 : 0;
    return kRegion ? w : -1;
}
static_assert(!kRegion || kPackStart, "the systolic kernel carries the START cell in one packed word");

// one cell's working set: every state's score + all possible words (dead ones fold away)
constexpr int CWMAX = 1 + NSH + 1;
constexpr int UNSET = INT_MIN;   // a state no transition has reached yet in this cell
// LOCAL models (START and END in scope everywhere: END is reachable with a score >= 0 from every
// cell) run RELAXED: only the transitions that leave START are tested for validity.  A transition
// whose source cell does not exist (row or column < 0) reads a register that holds <= LOWV (the
// reset value, viterbi.c:691-694) and offers LOWV-ish garbage, which no value that descends from
// START can lose against -- so every REACHABLE value, every winner on a path from START and the
// END maximum are exactly the reference's; only cells the reference leaves unset differ, and
// those are clamped back to LOWV when stored.  Models with restricted scopes (BSDP's derived
// models, global / bestfit / overlap), where an unreachable END must stay unset, keep every test.
constexpr bool kRelaxed = START_SCOPE == C4B_SCOPE_ANYWHERE && END_SCOPE == C4B_SCOPE_ANYWHERE;
struct Cell {
    int v[S][CWMAX];   // [0] score (UNSET until assigned), [1 + l] shadow slot l, [1 + NSH] packed start cell
    unsigned char win[S];
};

struct SysCtx {
    Ctx X;
    int row0;          // lattice row of real row 0 of this lane
    bool has_up;       // some lattice row lies above my strip
    int j;             // my column this step
    bool colok;        // 0 <= j <= T
    unsigned bmask;    // JIT_SYS_BLK: rows of my strip whose cell in this column is SubOpt-blocked
};

template <int K, int ROW>
__device__ __forceinline__ void sys_precalc(const SysCtx &Z, const int (&V)[NROWS][VW], int (&cs)[TN]) {
    if constexpr (K < TN) {
        if constexpr (calc_hoisted<K>()) {
            constexpr int calc = kTrCalc[K];
            constexpr int slot = calc_shadow_slot<calc>();
            constexpr int in = kTrIn[K], aq = kTrAq[K], at = kTrAt[K];
            const int si = max(Z.row0 + ROW - aq, 0), sj = max(Z.j - at, 0);
            int shadow = 0;
            if constexpr (slot >= 0 && word_of_slot(in, slot >= 0 ? slot : 0) >= 0)
                shadow = V[AQ + ROW - aq][kVOff[in] + at * kNW[in] + word_of_slot(in, slot >= 0 ? slot : 0)];
            cs[K] = calc_score<calc>(Z.X, Z.X.q_start + si, Z.X.t_start + min(sj, Z.X.T), shadow);
        }
        sys_precalc<K + 1, ROW>(Z, V, cs);
    }
}

template <int K, int ROW>
__device__ __forceinline__ void sys_transitions(const SysCtx &Z, const int (&V)[NROWS][VW], const int (&cs)[TN],
                                                bool rowok, Cell &c) {
    if constexpr (K < TN) {
        constexpr int in = kTrIn[K], out = kTrOut[K], aq = kTrAq[K], at = kTrAt[K];
        constexpr int calc = kTrCalc[K];
        constexpr bool from_start = (in == START);
        const int i = Z.row0 + ROW, j = Z.j;
        const int si = i - aq, sj = j - at;
        constexpr bool tested = !kRelaxed || from_start;
        bool valid = true;
        if constexpr (tested) {
            valid = rowok && state_active<in>(si, sj, Z.X.Q, Z.X.T) && state_active<out>(i, j, Z.X.Q, Z.X.T);
            if constexpr (at > 0) valid = valid && sj >= 0;
            if constexpr (aq > ROW) valid = valid && (Z.has_up && si >= 0);
        }
        int src[CWMAX];
#pragma unroll
        for (int l = 0; l < CWMAX; ++l) src[l] = 0;
        if constexpr (from_start) {
            // START's cell is all zero (no cell_start_func tables on this kernel)
        } else if constexpr (aq + at > 0) {
            constexpr int base = kVOff[in] + at * kNW[in];
            src[0] = V[AQ + ROW - aq][base];
#pragma unroll
            for (int l = 0; l < NSH; ++l)
                if (word_of_slot(in, l) >= 0) src[1 + l] = V[AQ + ROW - aq][base + max(word_of_slot(in, l), 0)];
            if constexpr (kRegion) src[1 + NSH] = V[AQ + ROW - aq][base + max(word_of_start(in), 0)];
        } else {
#pragma unroll
            for (int l = 0; l < CWMAX; ++l) src[l] = c.v[in][l];
        }
        int t = src[0];
        if constexpr (!from_start && aq + at == 0) {   // an unset state reads as reset (viterbi.c:691-694)
            if constexpr (kRelaxed) t = max(t, LOWV);
            else t = (t == UNSET) ? LOWV : t;
        }
        if constexpr (calc >= 0) {
            if constexpr (calc_hoisted<K>()) {
                t += cs[K];
            } else {  // shadow-reading calc on a silent transition: its slot is only known now
                constexpr int slot = calc_shadow_slot<calc>();
                t += calc_score<calc>(Z.X, Z.X.q_start + max(si, 0), Z.X.t_start + max(sj, 0), src[1 + (slot >= 0 ? slot : 0)]);
            }
            if constexpr ((kCalcProt[calc >= 0 ? calc : 0] & C4B_PROTECT_UNDERFLOW) != 0) t = max(t, LOWV);
            if constexpr ((kCalcProt[calc >= 0 ? calc : 0] & C4B_PROTECT_OVERFLOW) != 0)
                t = min(t, C4B_IMPOSSIBLY_HIGH_SCORE);
        }
        // "first valid transition assigns, later ones replace only if strictly greater"
        // (viterbi.c:766-775) as a max: an unset state holds UNSET, below every candidate, and an
        // invalid transition offers UNSET.  (t is never UNSET: scores stay above 2 * LOWV.)
        int teff = tested ? (valid ? t : UNSET) : t;
        if constexpr (kBlk && kTrLabel[K] == C4B_LABEL_MATCH)   // viterbi.c:701-704
            teff = ((Z.bmask >> ROW) & 1u) ? UNSET : teff;
        constexpr bool carries = (JIT_MODE == GEN_PATH) || kRegion || (kNW[out] > 1);
        if constexpr (carries) {
            const bool take = teff > c.v[out][0];
            // Viterbi_Data_assign (viterbi.c:445-462); stamps go on the transported copy
#pragma unroll
            for (int l = 0; l < NSH; ++l) {
                if (word_of_slot(out, l) >= 0) {   // (folds per unrolled l: dead slots cost nothing)
                    const int stamp = kShadow[in * C4B_MAX_SHADOW_SLOTS + l];
                    int nv = src[1 + l];
                    if (stamp == 1) nv = Z.X.t_start + sj;
                    if (stamp == 2) nv = Z.X.q_start + si;
                    c.v[out][1 + l] = take ? nv : c.v[out][1 + l];
                }
            }
            if constexpr (kRegion) {
                int nv = src[1 + NSH];
                if constexpr (from_start) nv = si * (Z.X.T + 1) + sj;
                c.v[out][1 + NSH] = take ? nv : c.v[out][1 + NSH];
            }
            if constexpr (JIT_MODE == GEN_PATH) c.win[out] = take ? (unsigned char)kTbCode[K] : c.win[out];
        }
        c.v[out][0] = max(c.v[out][0], teff);
        sys_transitions<K + 1, ROW>(Z, V, cs, rowok, c);
    }
}

// What a lane hands to the lane below per step (host-generated): word w of the hand-off is
// V[AQ + SR - kSendD[w]][kSendOff[w]] of the sender (current column of the row at distance
// kSendD[w] above the receiver's row 0) and lands in V[AQ - kSendD[w]][kSendOff[w]] of the
// receiver -- the carried words of every state some transition reads at advance_query >= d.

// ---- one lattice row of the lane's strip at the lane's column -----------------------
struct SysBest {
    int score, i, j, start;
};

template <int ROW>
__device__ __forceinline__ void sys_row(const SysCtx &Z, int (&V)[NROWS][VW], SysBest &best,
                                        uint32_t (&tbw)[TB_CHUNK / 4]) {
    const int i = Z.row0 + ROW;
    const bool rowok = Z.colok && i <= Z.X.Q;
    int cs[TN];
    sys_precalc<0, ROW>(Z, V, cs);
    Cell c;
#pragma unroll
    for (int st = 0; st < S; ++st) {
#pragma unroll
        for (int l = 0; l < CWMAX; ++l) c.v[st][l] = (l == 0) ? UNSET : 0;
        c.win[st] = 0;
    }
    sys_transitions<0, ROW>(Z, V, cs, rowok, c);
    {   // viterbi.c:778-791; within a thread cells arrive in scan order; UNSET never beats a score
        const int v = c.v[END][0];
        if ((!kRelaxed || rowok) && v > best.score) {
            best.score = v; best.i = i; best.j = Z.j;
            if constexpr (kRegion) best.start = c.v[END][1 + NSH];
        }
    }
    if constexpr (JIT_MODE == GEN_PATH) {
        // the record: per state the rank of the winner among the transitions entering it (0 = unset)
#pragma unroll
        for (int st = 0; st < S; ++st)
            if (kTbBits[st] > 0) {
                const int pos = ROW * TB_ROW_BITS + kTbBitOff[st];
                tbw[pos / 32] |= (uint32_t)c.win[st] << (pos % 32);
                if (pos % 32 + kTbBits[st] > 32) tbw[pos / 32 + 1] |= (uint32_t)c.win[st] >> (32 - pos % 32);
            }
    }
    // current column of the states that are read later (an unset state is stored as the reset value)
#pragma unroll
    for (int st = 0; st < S; ++st)
        if (kVD[st] >= 0) {
            V[AQ + ROW][kVOff[st]] = kRelaxed ? max(c.v[st][0], LOWV) : ((c.v[st][0] == UNSET) ? LOWV : c.v[st][0]);
#pragma unroll
            for (int l = 0; l < NSH; ++l)
                if (word_of_slot(st, l) >= 0) V[AQ + ROW][kVOff[st] + max(word_of_slot(st, l), 0)] = c.v[st][1 + l];
            if (kRegion) V[AQ + ROW][kVOff[st] + max(word_of_start(st), 0)] = c.v[st][1 + NSH];
        }
}

template <int ROW>
__device__ __forceinline__ void sys_rows(const SysCtx &Z, int (&V)[NROWS][VW], SysBest &best,
                                         uint32_t (&tbw)[TB_CHUNK / 4]) {
    if constexpr (ROW < SR) {   // top-down: a row reads the CURRENT column of the rows above it
        sys_row<ROW>(Z, V, best, tbw);
        sys_rows<ROW + 1>(Z, V, best, tbw);
    }
}

}  // namespace c4bjit

// one CTA per lattice; blockDim.x = 32 W, the W warps take the strips of 32 R rows round-robin
// JIT_SYS_WARPS = warps per CTA the host launches (strips in flight per lattice), JIT_SYS_MINB =
// resident CTAs per SM the register allocation must allow (the host aims at 16 warps per SM)
extern "C" __global__ void __launch_bounds__(32 * JIT_SYS_WARPS, JIT_SYS_MINB)
c4b_jit_sys(const c4b::GenPair *__restrict__ pairs, int n_pairs, c4b::GenOut *__restrict__ outs,
            const c4b::GenTables *__restrict__ tables, int32_t *top_base, size_t top_stride,
            const c4b::GenWin *__restrict__ wins) {
    using namespace c4bjit;
    __shared__ c4b_scoring s_scoring;
    __shared__ volatile long long vprog[kSysMaxWarps];
    __shared__ int red[kSysMaxWarps][4];
    {
        const int *src = reinterpret_cast<const int *>(&tables->scoring);
        int *dst = reinterpret_cast<int *>(&s_scoring);
        for (int k = threadIdx.x; k < (int)(sizeof(c4b_scoring) / 4); k += blockDim.x) dst[k] = src[k];
    }
    if (threadIdx.x < kSysMaxWarps) vprog[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int pi = blockIdx.x;
    if (pi >= n_pairs) return;
    const GenPair P = pairs[pi];
    SysCtx Z;
    Z.X.sc = &s_scoring;
    Z.X.q = P.q; Z.X.t = P.t;
    for (int k = 0; k < 4; ++k) Z.X.splice[k] = P.splice[k];
    Z.X.start_cells = nullptr;
    Z.X.blk_q = nullptr; Z.X.blk_t = nullptr;
    Z.X.n_blocked = 0; Z.X.blk_dq = 0; Z.X.blk_dt = 0;
    Z.X.q_start = P.q_start; Z.X.t_start = P.t_start; Z.X.Q = P.Q; Z.X.T = P.T;
    Z.bmask = 0u;
    const int Q = P.Q, T = P.T;
    constexpr int rows_per_sweep = 32 * SR;
    const int all_sweeps = (Q + 1 + rows_per_sweep - 1) / rows_per_sweep;
    int nsweeps = all_sweeps;
    const int nsteps = T + 1 + 31;
    const int pub_mask = (T >= 4096) ? 31 : 7;
    // column window [c0, c1] of this launch (the whole lattice unless JIT_SYS_WIN == 2)
    int c0 = 0, c1 = T, wcols = 0;
    int32_t *ck = nullptr;
    constexpr int CKW = NROWS * VW;   // words of one lane's checkpoint
    if constexpr (kWinMode != 0) {
        const c4b::GenWin Wn = wins[pi];
        ck = Wn.ck; wcols = Wn.wcols;
        if constexpr (kWinMode == 2) { c0 = Wn.c0; c1 = Wn.c1; nsweeps = min(all_sweeps, Wn.nsweeps); }
    }
    // sweep hand-off rows in L2: two buffers of (T + 1) x NSEND words per lattice, ping-pong
    int32_t *top0 = top_base + (size_t)pi * top_stride;
    int32_t *top1 = top0 + (size_t)(T + 1) * (NSEND > 0 ? NSEND : 1);

    SysBest bst;
    bst.score = INT_MIN; bst.i = 0; bst.j = 0; bst.start = 0;

    for (int sweep = warp; sweep < nsweeps; sweep += W) {
        Z.row0 = sweep * rows_per_sweep + lane * SR;
        Z.has_up = !(sweep == 0 && lane == 0);
        const bool later_sweep = sweep > 0;
        const int32_t *top_in = (sweep & 1) ? top0 : top1;   // written by sweep - 1
        int32_t *top_out = (sweep & 1) ? top1 : top0;
        const bool write_top = (sweep + 1 < nsweeps) && lane == 31;
        const bool piped = later_sweep && W > 1;
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
        // the register lattice: V[row][state block: column k = 0 (current) .. kVD][carried word]
        int V[NROWS][VW];
#pragma unroll
        for (int r = 0; r < NROWS; ++r)
#pragma unroll
            for (int st = 0; st < S; ++st)
                if (kVD[st] >= 0) {
#pragma unroll
                    for (int k = 0; k <= (kVD[st] >= 0 ? kVD[st] : 0); ++k)
#pragma unroll
                        for (int w = 0; w < kNW[st]; ++w) V[r][kVOff[st] + k * kNW[st] + w] = (w == 0) ? LOWV : 0;
                }
        // lane 0 of a later sweep: column 0 of the row(s) above comes from the hand-off buffer
        int upin[NSEND > 0 ? NSEND : 1];
#pragma unroll
        for (int w = 0; w < (NSEND > 0 ? NSEND : 1); ++w) upin[w] = LOWV;   // "no row above": reads as reset
        auto load_top = [&](int col) {
#pragma unroll
            for (int w = 0; w < NSEND; ++w) upin[w] = __ldcg(top_in + (size_t)col * NSEND + w);
        };
        auto place_up = [&](const int (&vals)[NSEND > 0 ? NSEND : 1]) {   // received words -> virtual rows, column 0
#pragma unroll
            for (int w = 0; w < NSEND; ++w) V[AQ - kSendD[w]][kSendOff[w]] = vals[w];
        };
        if constexpr (kWinMode == 2) {
            if (c0 > 0) {   // the state after column c0 - 1: [window boundary][sweep][word][lane]
                const int32_t *cp = ck + (((size_t)(c0 / wcols - 1) * all_sweeps + sweep) * CKW) * 32 + lane;
#pragma unroll
                for (int r = 0; r < NROWS; ++r)
#pragma unroll
                    for (int x = 0; x < VW; ++x) V[r][x] = cp[(size_t)(r * VW + x) * 32];
            }
        }
        if (later_sweep) {
            if (piped) wait_column(c0);
            load_top(c0);
            if (lane == 0) place_up(upin);
        }
        // SubOpt: cursor into my strip's blocked columns
        int blk_cur = 0, blk_end = 0, blk_next = 0x7fffffff;
        if constexpr (kBlk) {
            if (P.blk_off != nullptr) {
                blk_cur = P.blk_off[sweep * 32 + lane];
                blk_end = P.blk_off[sweep * 32 + lane + 1];
                if constexpr (kWinMode == 2) {   // first entry at or after the window's first column
                    int lo = blk_cur, hi = blk_end;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (P.blk[mid].x < c0) lo = mid + 1; else hi = mid;
                    }
                    blk_cur = lo;
                }
                if (blk_cur < blk_end) blk_next = P.blk[blk_cur].x;
            }
        }
        unsigned char *tbp = nullptr;
        constexpr int TBCH = TB_CHUNK;   // traceback bytes per lane per step
        if constexpr (JIT_MODE == GEN_PATH)
            tbp = P.tb + ((size_t)sweep * (kWinMode == 2 ? wcols + 31 : nsteps) * 32 + lane) * TBCH;

        for (int s = c0; s < c1 + 32; ++s) {
            const int j = s - lane;
            Z.j = j;
            Z.colok = (j >= 0 && j <= T);
            // a lane left of its window waits for its first column with the checkpoint untouched (the
            // first window has no checkpoint: its lanes run in from column -lane like a whole-lattice pass)
            const bool live = (kWinMode != 2) || c0 == 0 || (j >= c0);
            const bool inwin = (kWinMode != 2) || (j >= c0 && j <= c1);
            // the next column of the sweep above, fetched early with warp-uniform addresses
            // (c1 = T outside window refills; a refill's producer never goes past c1)
            if (later_sweep && s + 1 <= c1) {
                if (piped) wait_column(s + 1);
                load_top(s + 1);
            }
            uint32_t tbw[TBCH / 4];
#pragma unroll
            for (int k = 0; k < TBCH / 4; ++k) tbw[k] = 0u;
            if constexpr (kBlk) {
                Z.bmask = 0u;
                if (j == blk_next) {
                    Z.bmask = (unsigned)P.blk[blk_cur].y;
                    ++blk_cur;
                    blk_next = (blk_cur < blk_end) ? P.blk[blk_cur].x : 0x7fffffff;
                }
            }
            if (live) sys_rows<0>(Z, V, bst, tbw);

            if constexpr (JIT_MODE == GEN_PATH) {
                if (Z.colok && inwin) {
#pragma unroll
                    for (int k = 0; k < TBCH / 4; ++k) reinterpret_cast<uint32_t *>(tbp)[k] = tbw[k];
                }
                tbp += 32 * TBCH;
            }
            // hand the bottom rows' current column to the lane below / the next sweep
            int send[NSEND > 0 ? NSEND : 1];
#pragma unroll
            for (int w = 0; w < NSEND; ++w) send[w] = V[AQ + SR - kSendD[w]][kSendOff[w]];
            if (write_top && Z.colok && inwin) {
#pragma unroll
                for (int w = 0; w < NSEND; ++w) top_out[(size_t)j * NSEND + w] = send[w];
                // publish in groups of 32 (8 on short lattices) columns: the fence costs far more than the
                // stores, and the consumer runs at least 32 columns behind anyway
                if (W > 1 && ((j & pub_mask) == pub_mask || j == c1)) {
                    __threadfence_block();   // the rows are written before the counter moves
                    vprog[warp] = (long long)sweep * (T + 1) + j + 1;
                }
            }
            int recv[NSEND > 0 ? NSEND : 1];
#pragma unroll
            for (int w = 0; w < NSEND; ++w) {
                recv[w] = __shfl_up_sync(0xffffffffu, send[w], 1);
                if (lane == 0) recv[w] = upin[w];   // column s + 1 of the sweep above (unused in sweep 0)
            }
            // every kept column moves one place back; column 0 of the virtual rows is what arrived
            if (live) {
#pragma unroll
                for (int r = 0; r < NROWS; ++r)
#pragma unroll
                    for (int st = 0; st < S; ++st)
                        if (kVD[st] >= 1) {
#pragma unroll
                            for (int k = (kVD[st] >= 1 ? kVD[st] : 1); k >= 1; --k)
#pragma unroll
                                for (int w = 0; w < kNW[st]; ++w)
                                    V[r][kVOff[st] + k * kNW[st] + w] = V[r][kVOff[st] + (k - 1) * kNW[st] + w];
                        }
                place_up(recv);
            }
            if constexpr (kWinMode == 1) {
                // last column of a window (wcols is a power of two): V now holds everything column j + 1
                // of my rows depends on, the hand-off from the lane above included
                if (Z.colok && j < T && ((j + 1) & (wcols - 1)) == 0) {
                    int32_t *cp = ck + (((size_t)((j + 1) / wcols - 1) * all_sweeps + sweep) * CKW) * 32 + lane;
#pragma unroll
                    for (int r = 0; r < NROWS; ++r)
#pragma unroll
                        for (int x = 0; x < VW; ++x) cp[(size_t)(r * VW + x) * 32] = V[r][x];
                }
            }
        }
        __syncwarp();
    }
    if constexpr (kWinMode == 2) return;   // (END is known: the cursor came from pass 1)
    int best = bst.score, best_i = bst.i, best_j = bst.j, best_start = bst.start;
    // lexicographic reduction: max score, then min j, then min i
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
        const int os = __shfl_xor_sync(0xffffffffu, best_start, off);
        const bool take = (ob > best) || (ob == best && (oj < best_j || (oj == best_j && oi < best_i)));
        if (take) { best = ob; best_j = oj; best_i = oi; best_start = os; }
    }
    if (W > 1) {
        if (lane == 0) { red[warp][0] = best; red[warp][1] = best_j; red[warp][2] = best_i; red[warp][3] = best_start; }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 1; w < W; ++w) {
                const int ob = red[w][0], oj = red[w][1], oi = red[w][2];
                if ((ob > best) || (ob == best && (oj < best_j || (oj == best_j && oi < best_i)))) {
                    best = ob; best_j = oj; best_i = oi; best_start = red[w][3];
                }
            }
    }
    if (threadIdx.x == 0) {
        GenOut o;
        o.score = best; o.end_i = best_i; o.end_j = best_j;
        o.start_i = kRegion ? best_start / (T + 1) : 0;
        o.start_j = kRegion ? best_start % (T + 1) : 0;
        o.flags = (best == INT_MIN) ? 1 : 0;
        outs[P.out_index] = o;
    }
}
