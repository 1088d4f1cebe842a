// generic_jit_kernel.cuh -- the table-driven lattice fill, SPECIALISED to one
// closed C4 model at run time (NVRTC; never seen by nvcc).
//
// The reference gets its speed on the CPU by generating C for each model and
// compiling it into the binary (src/c4/codegen.c, viterbi.c:1638-1727,
// bootstrapper.c).  This is the device form of the same idea: the host
// (generic_jit.inl) writes the closed model out as constexpr tables
// (c4bjit::kTr*, kCalc*, kShadow, ...) in front of this file and compiles it
// for sm_100a.  Every loop over transitions / states / shadow slots below is a
// compile-time recursion, so
//   * a cell's states live in registers (cur[S*C]) instead of a global ring,
//   * validity tests, calc kinds, shadow stamps and protect flags fold away,
//   * all lattice loads of a cell (sources with a non-zero advance) are issued
//     together before the max-plus chain instead of one dependent round trip
//     per transition,
//   * only states that feed a non-silent transition are written back, each
//     for just as many columns as its longest advance needs, in a ring laid
//     out [state][column slot][word][row] so a diagonal's threads touch
//     consecutive words -- small enough to live in shared memory for queries
//     up to several hundred rows (JIT_SMEM_RING), else in L2.
// Semantics are those of Viterbi_interpreted (src/c4/viterbi.c:655-837), cell
// for cell the same as generic_wavefront.cuh (the interpreter kernel, which is
// the reading reference for this file).
//
// Expected in front of this file: JIT_MODE (GEN_*), JIT_THREADS, JIT_MIN_CTAS, JIT_SMEM_RING,
// JIT_PACK_START, and namespace
// c4bjit { S, TN, NSH, START, END, START_SCOPE, END_SCOPE,
// kTrIn/kTrOut/kTrAq/kTrAt/kTrCalc/kTrLabel[TN], kCalcKind/kCalcProt/kCalcP0/
// kCalcP1/kCalcP2[], kShadow[S*C4B_MAX_SHADOW_SLOTS], kDepth[S], kRingOff[S] }.

namespace c4bjit {
using namespace c4b;

constexpr int LOWV = C4B_IMPOSSIBLY_LOW_SCORE;
constexpr bool kRegion = (JIT_MODE == GEN_REGION) && START_SCOPE != C4B_SCOPE_CORNER;
// REGION mode carries where the path left START in extra cell slots (viterbi.c:403-411):
// QID / TID, or -- when the host saw that every lattice of the batch has fewer than 2^31
// cells (JIT_PACK_START) -- both in the one slot QID as start_i * (T+1) + start_j.
constexpr bool kPackStart = JIT_PACK_START && kRegion && START_SCOPE != C4B_SCOPE_QUERY &&
                            START_SCOPE != C4B_SCOPE_TARGET;
constexpr int QID = (kRegion && START_SCOPE != C4B_SCOPE_QUERY) ? 1 + NSH : -1;
constexpr int TID = (kRegion && START_SCOPE != C4B_SCOPE_TARGET && !kPackStart) ? 1 + NSH + (QID >= 0 ? 1 : 0) : -1;
constexpr int C = 1 + NSH + (QID >= 0 ? 1 : 0) + (TID >= 0 ? 1 : 0);
// The lattice ring keeps, per saved state s, only the kDepth[s] most recent
// columns a reader can still ask for (largest advance_query + advance_target
// of the transitions leaving s, plus one): state s owns ring rows
// [kRingOff[s], kRingOff[s] + kDepth[s]) x C words x pitch, column j in row
// j % kDepth[s].  (Cell (i,j) is last read on diagonal i+j+max advance and
// overwritten by column j + kDepth[s] on diagonal i+j+kDepth[s].)

struct Ctx {
    const c4b_scoring *sc;  // shared-memory copy
    const uint8_t *q, *t;
    const int32_t *splice[4];
    const int32_t *start_cells;
    const int32_t *blk_q, *blk_t;
    int n_blocked, blk_dq, blk_dt;
    int q_start, t_start, Q, T;
};

// Layout_is_transition_valid, one state at one cell (src/c4/layout.c:21-88)
template <int STATE>
__device__ __forceinline__ bool state_active(int qp, int tp, int Q, int T) {
    bool ok = true;
    if constexpr (STATE == START) {
        if constexpr (START_SCOPE == C4B_SCOPE_EDGE) ok = ok && (qp == 0 || tp == 0);
        if constexpr (START_SCOPE == C4B_SCOPE_QUERY) ok = ok && qp == 0;
        if constexpr (START_SCOPE == C4B_SCOPE_TARGET) ok = ok && tp == 0;
        if constexpr (START_SCOPE == C4B_SCOPE_CORNER) ok = ok && qp == 0 && tp == 0;
    }
    if constexpr (STATE == END) {
        if constexpr (END_SCOPE == C4B_SCOPE_EDGE) ok = ok && (qp == Q || tp == T);
        if constexpr (END_SCOPE == C4B_SCOPE_QUERY) ok = ok && qp == Q;
        if constexpr (END_SCOPE == C4B_SCOPE_TARGET) ok = ok && tp == T;
        if constexpr (END_SCOPE == C4B_SCOPE_CORNER) ok = ok && qp == Q && tp == T;
    }
    return ok;
}

__device__ __forceinline__ int submat(const int32_t *matrix, const uint8_t *index, int a, int b) {
    const int ia = index[a & 255], ib = index[b & 255];
    return matrix[min(ia, 23) * C4B_SUBMAT_N + min(ib, 23)];
}
__device__ __forceinline__ int translate(const c4b_scoring &s, int a, int b, int c) {
    return s.codon_aa[s.nt2d[a & 255] | (s.nt2d[b & 255] << 4) | (s.nt2d[c & 255] << 8)];
}

// C4_Calc_score + the calc callbacks (include/c4b200.h); `shadow` = slot
// kCalcP2 of the SOURCE cell where the kind reads one.  Branch-free: every
// lookup is issued (at a clamped index where the reference would not look) and
// the -infinity cases are selects, so calcs of one cell overlap.
#define SEQ(p, k) ((int)__ldg((p) + (k)))
template <int CALC>
__device__ __forceinline__ int calc_score(const Ctx &X, int qp, int tp, int shadow) {
    if constexpr (CALC < 0) {
        return 0;
    } else {
        constexpr int kind = kCalcKind[CALC];
        constexpr int p0 = kCalcP0[CALC], p1 = kCalcP1[CALC];
        const c4b_scoring &s = *X.sc;
        const uint8_t *q = X.q, *t = X.t;
        if constexpr (kind == C4B_CALC_CONST) return p0;
        else if constexpr (kind == C4B_CALC_MATCH_DNA) return submat(s.dna_matrix, s.dna_index, SEQ(q, qp), SEQ(t, tp));
        else if constexpr (kind == C4B_CALC_MATCH_PROTEIN)
            return submat(s.protein_matrix, s.protein_index, SEQ(q, qp), SEQ(t, tp));
        else if constexpr (kind == C4B_CALC_MATCH_1_3)
            return submat(s.protein_matrix, s.protein_index, SEQ(q, qp),
                          translate(s, SEQ(t, tp), SEQ(t, tp + 1), SEQ(t, tp + 2)));
        else if constexpr (kind == C4B_CALC_MATCH_3_1)
            return submat(s.protein_matrix, s.protein_index, translate(s, SEQ(q, qp), SEQ(q, qp + 1), SEQ(q, qp + 2)),
                          SEQ(t, tp));
        else if constexpr (kind == C4B_CALC_MATCH_3_3)
            return submat(s.protein_matrix, s.protein_index, translate(s, SEQ(q, qp), SEQ(q, qp + 1), SEQ(q, qp + 2)),
                          translate(s, SEQ(t, tp), SEQ(t, tp + 1), SEQ(t, tp + 2)));
        else if constexpr (kind == C4B_CALC_SPLICE_PRE) return p0 + __ldg(X.splice[p1] + tp);
        else if constexpr (kind == C4B_CALC_SPLICE_POST) {
            // min_intron <= len <= max_intron (intron.c:151-159) as ONE unsigned compare
            const int len = tp - shadow + 2;
            const int v = __ldg(X.splice[p1] + tp);
            const bool bad = (s.max_intron < s.min_intron) |
                             ((unsigned)(len - s.min_intron) > (unsigned)(s.max_intron - s.min_intron));
            return bad ? LOWV : v;
        } else if constexpr (kind == C4B_CALC_PHASE1_POST) {
            const int sh = max(shadow, 1);
            const int v = submat(s.protein_matrix, s.protein_index, SEQ(q, qp),
                                 translate(s, SEQ(t, sh - 1), SEQ(t, tp), SEQ(t, tp + 1)));
            return shadow < 1 ? LOWV : v;
        } else if constexpr (kind == C4B_CALC_PHASE2_POST) {
            const int sh = max(shadow, 2);
            const int v = submat(s.protein_matrix, s.protein_index, SEQ(q, qp),
                                 translate(s, SEQ(t, sh - 2), SEQ(t, sh - 1), SEQ(t, tp)));
            return shadow < 2 ? LOWV : v;
        } else {
            return LOWV;
        }
    }
}

template <int CALC>
__device__ constexpr int calc_shadow_slot() {
    if constexpr (CALC < 0) return -1;
    else return (kCalcKind[CALC] >= C4B_CALC_SPLICE_POST) ? kCalcP2[CALC] : -1;
}

__device__ __forceinline__ bool blocked(const Ctx &X, int i, int j) {
    // SubOpt_Index lookup (src/c4/subopt.c:250-374) as an exact set
    const int bi = i + X.blk_dq, bj = j + X.blk_dt;
    int lo = 0, hi = X.n_blocked;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int tj = __ldg(X.blk_t + mid), qi = __ldg(X.blk_q + mid);
        if (tj < bj || (tj == bj && qi < bi)) lo = mid + 1;
        else hi = mid;
    }
    return lo < X.n_blocked && __ldg(X.blk_t + lo) == bj && __ldg(X.blk_q + lo) == bi;
}

__device__ constexpr bool tr_from_ring(int k) { return kTrIn[k] != START && (kTrAq[k] + kTrAt[k]) > 0; }

// ---- phase 1: every lattice load of the cell, issued back to back -----------
template <int K>
__device__ __forceinline__ void preload(const int32_t *ring, int pitch, int i, int j, int (&pre)[TN * C]) {
    if constexpr (K < TN) {
        if constexpr (tr_from_ring(K)) {
            constexpr int aq = kTrAq[K], at = kTrAt[K], in = kTrIn[K];
            const int si = i - aq, sj = j - at;
            const bool ok = si >= 0 && sj >= 0;
            const int sslot = ok ? sj % kDepth[in] : 0;
            const int32_t *src = ring + ((kRingOff[in] + sslot) * C) * pitch + si;
#pragma unroll
            for (int l = 0; l < C; ++l) pre[K * C + l] = ok ? src[l * pitch] : 0;
        }
        preload<K + 1>(ring, pitch, i, j, pre);
    }
}

// ---- phase 1b: every calc whose inputs are known before the chain -------------
// (all but shadow-reading calcs on silent transitions).  Evaluated for every
// transition whether or not it is valid at this cell -- coordinates are clamped
// into the staged buffers -- so the sequence / splice / matrix lookups of one
// cell overlap instead of queueing behind 20-odd branches.
template <int K>
__device__ constexpr bool calc_hoisted() {
    return kTrCalc[K] >= 0 && (calc_shadow_slot<kTrCalc[K]>() < 0 || tr_from_ring(K));
}

template <int K>
__device__ __forceinline__ void precalc(const Ctx &X, int i, int j, const int (&pre)[TN * C], int (&cs)[TN]) {
    if constexpr (K < TN) {
        if constexpr (calc_hoisted<K>()) {
            constexpr int calc = kTrCalc[K];
            constexpr int slot = calc_shadow_slot<calc>();
            const int si = max(i - kTrAq[K], 0), sj = max(j - kTrAt[K], 0);
            int shadow = 0;
            if constexpr (slot >= 0) shadow = pre[K * C + 1 + (slot >= 0 ? slot : 0)];
            cs[K] = calc_score<calc>(X, X.q_start + si, X.t_start + sj, shadow);
        }
        precalc<K + 1>(X, i, j, pre, cs);
    }
}

// ---- phase 2: the transitions in closed-model order (viterbi.c:695-776) ------
template <int K>
__device__ __forceinline__ void transitions(const Ctx &X, int i, int j, bool match_blocked,
                                            const int (&pre)[TN * C], const int (&cs)[TN], int (&cur)[S * C],
                                            unsigned &set, unsigned char (&win)[S]) {
    if constexpr (K < TN) {
        constexpr int in = kTrIn[K], out = kTrOut[K], aq = kTrAq[K], at = kTrAt[K];
        constexpr int calc = kTrCalc[K];
        constexpr bool from_start = (in == START);
        const int si = i - aq, sj = j - at;
        bool valid = si >= 0 && sj >= 0 && state_active<in>(si, sj, X.Q, X.T) && state_active<out>(i, j, X.Q, X.T);
        if constexpr (kTrLabel[K] == C4B_LABEL_MATCH) valid = valid && !match_blocked;
        int src[C];
        if constexpr (from_start) {
#pragma unroll
            for (int l = 0; l < C; ++l) src[l] = 0;
            if (valid && X.start_cells != nullptr) {  // cell_start_func table (viterbi.c:727-741)
                const int32_t *scell = X.start_cells + ((size_t)si * (X.T + 1) + sj) * (1 + NSH);
#pragma unroll
                for (int l = 0; l <= NSH; ++l) src[l] = __ldg(scell + l);
            }
        } else if constexpr (aq + at > 0) {
#pragma unroll
            for (int l = 0; l < C; ++l) src[l] = pre[K * C + l];
        } else {
#pragma unroll
            for (int l = 0; l < C; ++l) src[l] = cur[in * C + l];
        }
        int t = src[0];
        if constexpr (calc >= 0) {
            if constexpr (calc_hoisted<K>()) {
                t += cs[K];
            } else {  // shadow-reading calc on a silent transition: its slot is only known now
                constexpr int slot = calc_shadow_slot<calc>();
                if (valid) t += calc_score<calc>(X, X.q_start + si, X.t_start + sj, src[1 + (slot >= 0 ? slot : 0)]);
            }
            if constexpr ((kCalcProt[calc >= 0 ? calc : 0] & C4B_PROTECT_UNDERFLOW) != 0) t = max(t, LOWV);
            if constexpr ((kCalcProt[calc >= 0 ? calc : 0] & C4B_PROTECT_OVERFLOW) != 0)
                t = min(t, C4B_IMPOSSIBLY_HIGH_SCORE);
        }
        if (valid && (!((set >> out) & 1u) || cur[out * C] < t)) {
            set |= 1u << out;
            // Viterbi_Data_assign (viterbi.c:445-462); stamps go on the transported copy
            cur[out * C] = t;
#pragma unroll
            for (int l = 1; l < C; ++l) cur[out * C + l] = src[l];
#pragma unroll
            for (int l = 0; l < NSH; ++l) {
                constexpr int base = in * C4B_MAX_SHADOW_SLOTS;
                if (kShadow[base + l] == 1) cur[out * C + 1 + l] = X.t_start + sj;
                else if (kShadow[base + l] == 2) cur[out * C + 1 + l] = X.q_start + si;
            }
            if constexpr (from_start) {
                if constexpr (QID >= 0) cur[out * C + (QID >= 0 ? QID : 0)] = kPackStart ? si * (X.T + 1) + sj : si;
                if constexpr (TID >= 0) cur[out * C + (TID >= 0 ? TID : 0)] = sj;
            }
            win[out] = (unsigned char)K;
        }
        transitions<K + 1>(X, i, j, match_blocked, pre, cs, cur, set, win);
    }
}

__device__ constexpr bool model_has_match_label() {
    for (int k = 0; k < TN; ++k)
        if (kTrLabel[k] == C4B_LABEL_MATCH) return true;
    return false;
}

}  // namespace c4bjit

#ifndef JIT_SYSTOLIC   // generic_jit_systolic.cuh follows instead and defines its own kernel
// grid = resident CTAs; each loops over lattices through an atomic cursor.
extern "C" __global__ void __launch_bounds__(JIT_THREADS, JIT_MIN_CTAS)
c4b_jit_fill(const c4b::GenPair *__restrict__ pairs, int n_pairs, c4b::GenOut *__restrict__ outs,
             const c4b::GenTables *__restrict__ tables, int32_t *ring_base, size_t ring_stride,
             int *__restrict__ cursor) {
    using namespace c4bjit;
    __shared__ c4b_scoring s_scoring;
    __shared__ int s_pair;
    __shared__ int red_score[JIT_THREADS], red_i[JIT_THREADS], red_j[JIT_THREADS];
    __shared__ int red_si[JIT_THREADS], red_sj[JIT_THREADS];
    {
        const int *src = reinterpret_cast<const int *>(&tables->scoring);
        int *dst = reinterpret_cast<int *>(&s_scoring);
        for (int k = threadIdx.x; k < (int)(sizeof(c4b_scoring) / 4); k += JIT_THREADS) dst[k] = src[k];
    }
    __syncthreads();
    // the lattice ring: shared memory when DEPTH columns of the longest query fit
    // (the host decides, JIT_SMEM_RING), else this CTA's slice of the global ring
#if JIT_SMEM_RING
    extern __shared__ int32_t s_ring[];
    int32_t *ring = s_ring;
#else
    int32_t *ring = ring_base + (size_t)blockIdx.x * ring_stride;
#endif

    for (;;) {
        if (threadIdx.x == 0) s_pair = atomicAdd(cursor, 1);
        __syncthreads();
        const int pi = s_pair;
        if (pi >= n_pairs) break;
        const GenPair P = pairs[pi];
        Ctx X;
        X.sc = &s_scoring;
        X.q = P.q; X.t = P.t;
        for (int k = 0; k < 4; ++k) X.splice[k] = P.splice[k];
        X.start_cells = P.start_cells;
        X.blk_q = P.blk_q; X.blk_t = P.blk_t;
        X.n_blocked = P.n_blocked; X.blk_dq = P.blk_dq; X.blk_dt = P.blk_dt;
        X.q_start = P.q_start; X.t_start = P.t_start; X.Q = P.Q; X.T = P.T;
        const int Q = P.Q, T = P.T;
        const int pitch = Q + 1;
        int best = INT_MIN, best_i = 0, best_j = 0, best_si = 0, best_sj = 0;
        for (int d = 0; d <= Q + T; ++d) {
            const int hi = min(Q, d);
            for (int i = (int)threadIdx.x; i <= hi; i += JIT_THREADS) {
                const int j = d - i;
                if (j > T) continue;
                int pre[TN * C];
                preload<0>(ring, pitch, i, j, pre);
                bool match_blocked = false;
                if constexpr (model_has_match_label())
                    if (X.n_blocked) match_blocked = blocked(X, i, j);
                int cur[S * C];
#pragma unroll
                for (int k = 0; k < S * C; ++k) cur[k] = (k % C == 0) ? LOWV : 0;  // viterbi.c:691-694
                unsigned set = 0;
                unsigned char win[S];
#pragma unroll
                for (int k = 0; k < S; ++k) win[k] = 0xFF;
                int cs[TN];
                precalc<0>(X, i, j, pre, cs);
                transitions<0>(X, i, j, match_blocked, pre, cs, cur, set, win);
                if ((set >> END) & 1u) {  // viterbi.c:778-791
                    const int v = cur[END * C];
                    if (P.end_cells) {  // cell_end_func input (viterbi.c:792-797), consumed by the host binding
#pragma unroll
                        for (int l = 0; l <= NSH; ++l)
                            P.end_cells[((size_t)i * (T + 1) + j) * (1 + NSH) + l] = cur[END * C + l];
                    }
                    if (v > best || (v == best && (j < best_j || (j == best_j && i < best_i)))) {
                        best = v; best_i = i; best_j = j;
                        if constexpr (QID >= 0) best_si = cur[END * C + (QID >= 0 ? QID : 0)];
                        if constexpr (TID >= 0) best_sj = cur[END * C + (TID >= 0 ? TID : 0)];
                    }
                }
                if constexpr (JIT_MODE == GEN_PATH) {
                    unsigned char *tbc = P.tb + GEN_TB_CELL(i, j, Q, S);
#pragma unroll
                    for (int k = 0; k < S; ++k)
                        if (win[k] != 0xFF) tbc[k] = win[k];
                }
                // states that feed a non-silent transition go back to the lattice ring
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (kDepth[s] > 0) {
                        int32_t *cell = ring + ((kRingOff[s] + j % (kDepth[s] > 0 ? kDepth[s] : 1)) * C) * pitch + i;
#pragma unroll
                        for (int l = 0; l < C; ++l) cell[l * pitch] = cur[s * C + l];
                    }
            }
            __syncthreads();
        }
        red_score[threadIdx.x] = best; red_i[threadIdx.x] = best_i; red_j[threadIdx.x] = best_j;
        red_si[threadIdx.x] = best_si; red_sj[threadIdx.x] = best_sj;
        __syncthreads();
        for (int off = JIT_THREADS / 2; off > 0; off >>= 1) {
            if ((int)threadIdx.x < off) {
                const int a = threadIdx.x, b = threadIdx.x + off;
                const bool take = red_score[b] > red_score[a] ||
                                  (red_score[b] == red_score[a] &&
                                   (red_j[b] < red_j[a] || (red_j[b] == red_j[a] && red_i[b] < red_i[a])));
                if (take) {
                    red_score[a] = red_score[b]; red_i[a] = red_i[b]; red_j[a] = red_j[b];
                    red_si[a] = red_si[b]; red_sj[a] = red_sj[b];
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            GenOut o;
            o.score = red_score[0]; o.end_i = red_i[0]; o.end_j = red_j[0];
            o.start_i = kPackStart ? red_si[0] / (T + 1) : red_si[0];
            o.start_j = kPackStart ? red_si[0] % (T + 1) : red_sj[0];
            o.flags = (red_score[0] == INT_MIN) ? 1 : 0;
            outs[P.out_index] = o;
        }
        __syncthreads();
    }
}
#endif  // JIT_SYSTOLIC
