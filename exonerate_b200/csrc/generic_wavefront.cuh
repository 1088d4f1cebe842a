// generic_wavefront.cuh -- table-driven lattice fill for ANY closed C4 model.
//
// Device counterpart of Viterbi_interpreted (src/c4/viterbi.c:655-837): the
// model's states / transitions / calcs are data (c4b_model), evaluated per cell
// in closed-model order with the reference's "first valid assigns, later
// replace only if strictly greater" rule.  One CTA per lattice walks the
// anti-diagonals (cells of a diagonal are independent: every transition
// either stays in the cell or comes from a strictly earlier diagonal); threads
// take cells of the diagonal.  State rows live in a per-CTA ring of
// max_target_advance + max_query_advance + 1 lattice columns.
//
// Path flow follows Optimal_find_path (src/c4/optimal.c:368-413): a REGION pass
// carries the start coordinates as two extra shadow slots (viterbi.c:403-411),
// then a PATH pass records the winning transition per state only inside the
// alignment's bounding box, and a walk (viterbi.c:342-392) emits the path.
//
// This is the coverage path (est2genome, protein2genome, coding2coding, SubOpt
// blocking ...).  The affine family takes affine_systolic.cuh instead.
#pragma once
#include <algorithm>
#include <map>
#include <vector>

#include "c4b_common.cuh"
#include "generic_types.h"

namespace c4b {

constexpr int kGenThreads = 1024;  // upper bound; the host launches 256 / 512 / 1024 by the longest query

__device__ __forceinline__ bool gen_state_active(const c4b_model &m, int state, int qp, int tp, int ql, int tl) {
    if (qp < 0 || tp < 0 || qp > ql || tp > tl) return false;
    if (state == m.start_state) {
        const int s = m.start_scope;
        if (s == C4B_SCOPE_EDGE && qp != 0 && tp != 0) return false;
        if (s == C4B_SCOPE_QUERY && qp != 0) return false;
        if (s == C4B_SCOPE_TARGET && tp != 0) return false;
        if (s == C4B_SCOPE_CORNER && (qp != 0 || tp != 0)) return false;
    }
    if (state == m.end_state) {
        const int s = m.end_scope;
        if (s == C4B_SCOPE_EDGE && qp != ql && tp != tl) return false;
        if (s == C4B_SCOPE_QUERY && qp != ql) return false;
        if (s == C4B_SCOPE_TARGET && tp != tl) return false;
        if (s == C4B_SCOPE_CORNER && (qp != ql || tp != tl)) return false;
    }
    return true;
}

__device__ __forceinline__ int gen_submat(const int32_t *matrix, const uint8_t *index, int a, int b) {
    const int ia = index[a & 255], ib = index[b & 255];
    // symbols outside the matrix alphabet are rejected on the host; clamp anyway
    return matrix[min(ia, 23) * C4B_SUBMAT_N + min(ib, 23)];
}
__device__ __forceinline__ int gen_translate(const c4b_scoring &s, int a, int b, int c) {
    return s.codon_aa[s.nt2d[a & 255] | (s.nt2d[b & 255] << 4) | (s.nt2d[c & 255] << 8)];
}

// C4_Calc_score + the reference's calc callbacks (see include/c4b200.h)
__device__ __forceinline__ int gen_calc(const GenTables &G, const GenPair &P, int calc_id, int qp, int tp,
                                        const int32_t *src) {
    if (calc_id < 0) return 0;
    const c4b_calc &c = G.model.calcs[calc_id];
    const c4b_scoring &s = G.scoring;
    const uint8_t *q = P.q, *t = P.t;
    switch (c.kind) {
    case C4B_CALC_CONST: return c.param[0];
    case C4B_CALC_MATCH_DNA: return gen_submat(s.dna_matrix, s.dna_index, q[qp], t[tp]);
    case C4B_CALC_MATCH_PROTEIN: return gen_submat(s.protein_matrix, s.protein_index, q[qp], t[tp]);
    case C4B_CALC_MATCH_1_3:
        return gen_submat(s.protein_matrix, s.protein_index, q[qp], gen_translate(s, t[tp], t[tp + 1], t[tp + 2]));
    case C4B_CALC_MATCH_3_1:
        return gen_submat(s.protein_matrix, s.protein_index, gen_translate(s, q[qp], q[qp + 1], q[qp + 2]), t[tp]);
    case C4B_CALC_MATCH_3_3:
        return gen_submat(s.protein_matrix, s.protein_index, gen_translate(s, q[qp], q[qp + 1], q[qp + 2]),
                          gen_translate(s, t[tp], t[tp + 1], t[tp + 2]));
    case C4B_CALC_SPLICE_PRE: return c.param[0] + P.splice[c.param[1]][tp];
    case C4B_CALC_SPLICE_POST: {
        const int len = tp - src[1 + c.param[2]] + 2;
        if (len < s.min_intron || len > s.max_intron) return LOW;
        return P.splice[c.param[1]][tp];
    }
    case C4B_CALC_PHASE1_POST: {
        const int slot = src[1 + c.param[2]];
        if (slot < 1) return LOW;
        return gen_submat(s.protein_matrix, s.protein_index, q[qp], gen_translate(s, t[slot - 1], t[tp], t[tp + 1]));
    }
    case C4B_CALC_PHASE2_POST: {
        const int slot = src[1 + c.param[2]];
        if (slot < 2) return LOW;
        return gen_submat(s.protein_matrix, s.protein_index, q[qp], gen_translate(s, t[slot - 2], t[slot - 1], t[tp]));
    }
    default: return LOW;
    }
}

__device__ __forceinline__ bool gen_blocked(const GenPair &P, int i, int j) {
    // SubOpt_Index lookup (src/c4/subopt.c:250-374) as an exact set
    const int bi = i + P.blk_dq, bj = j + P.blk_dt;
    int lo = 0, hi = P.n_blocked;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int tj = P.blk_t[mid], qi = P.blk_q[mid];
        if (tj < bj || (tj == bj && qi < bi)) lo = mid + 1;
        else hi = mid;
    }
    return lo < P.n_blocked && P.blk_t[lo] == bj && P.blk_q[lo] == bi;
}

// grid = resident CTAs; each loops over lattices through an atomic cursor.
__global__ void __launch_bounds__(kGenThreads)
generic_fill_kernel(const GenPair *__restrict__ pairs, int n_pairs, GenOut *__restrict__ outs,
                    const GenTables *__restrict__ tables, int mode, int32_t *ring_base,
                    size_t ring_stride, int *__restrict__ cursor, int smem_ring) {
    extern __shared__ int32_t s_ring[];  // the lattice ring of small lattices (BSDP region fills) stays on the SM
    __shared__ GenTables G;
    __shared__ int s_pair;
    __shared__ int red_score[kGenThreads], red_i[kGenThreads], red_j[kGenThreads];
    __shared__ int red_si[kGenThreads], red_sj[kGenThreads];
    {
        const int *src = reinterpret_cast<const int *>(tables);
        int *dst = reinterpret_cast<int *>(&G);
        for (int k = threadIdx.x; k < (int)(sizeof(GenTables) / 4); k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    const c4b_model &m = G.model;
    const int S = m.n_states, Tn = m.n_transitions;
    int C = 1 + m.n_shadow_slots, qid = -1, tid = -1;
    if (mode == GEN_REGION && m.start_scope != C4B_SCOPE_CORNER) {
        if (m.start_scope != C4B_SCOPE_QUERY) qid = C++;
        if (m.start_scope != C4B_SCOPE_TARGET) tid = C++;
    }
    const int depth = m.max_target_advance + m.max_query_advance + 1;
    int32_t *ring = smem_ring ? s_ring : ring_base + (size_t)blockIdx.x * ring_stride;

    for (;;) {
        if (threadIdx.x == 0) s_pair = atomicAdd(cursor, 1);
        __syncthreads();
        const int pi = s_pair;
        if (pi >= n_pairs) break;
        const GenPair P = pairs[pi];
        const int Q = P.Q, T = P.T;
        const size_t col_stride = (size_t)(Q + 1) * S * C;
        // stale shadow slots must at least be valid coordinates: zero the ring
        for (size_t k = threadIdx.x; k < col_stride * depth; k += blockDim.x) ring[k] = 0;
        __syncthreads();
        int best = INT32_MIN, best_i = 0, best_j = 0, best_si = 0, best_sj = 0;
        for (int d = 0; d <= Q + T; ++d) {
            const int lo = max(0, d - T), hi = min(Q, d);
            for (int i = lo + (int)threadIdx.x; i <= hi; i += blockDim.x) {
                const int j = d - i;
                const int jslot = j % depth;   // ring column of this cell; sources are jslot - advance (mod depth)
                int32_t *cell = ring + (size_t)jslot * col_stride + (size_t)i * S * C;
                for (int k = 0; k < S; ++k) cell[k * C] = LOW;  // viterbi.c:691-694
                uint32_t set = 0;
                for (int k = 0; k < Tn; ++k) {
                    const c4b_transition tr = m.transitions[k];
                    const int si = i - tr.advance_query, sj = j - tr.advance_target;
                    if (!gen_state_active(m, tr.input, si, sj, Q, T) ||
                        !gen_state_active(m, tr.output, i, j, Q, T))
                        continue;  // Layout_is_transition_valid
                    if (tr.label == C4B_LABEL_MATCH && P.n_blocked && gen_blocked(P, i, j)) continue;
                    int sslot = jslot - tr.advance_target;
                    if (sslot < 0) sslot += depth;
                    const int32_t *src = ring + (size_t)sslot * col_stride + ((size_t)si * S + tr.input) * C;
                    int32_t *dst = cell + tr.output * C;
                    const bool from_start = (tr.input == m.start_state);
                    const bool start_cb = from_start && P.start_cells;
                    if (start_cb) src = P.start_cells + ((size_t)si * (T + 1) + sj) * (1 + m.n_shadow_slots);
                    int t = from_start ? (start_cb ? src[0] : 0) : src[0];
                    t += gen_calc(G, P, tr.calc, P.q_start + si, P.t_start + sj, src);
                    if (tr.calc >= 0) {
                        const int prot = m.calcs[tr.calc].protect;
                        if ((prot & C4B_PROTECT_UNDERFLOW) && t < LOW) t = LOW;
                        if ((prot & C4B_PROTECT_OVERFLOW) && t > C4B_IMPOSSIBLY_HIGH_SCORE)
                            t = C4B_IMPOSSIBLY_HIGH_SCORE;
                    }
                    if ((set >> tr.output) & 1u) {
                        if (!(dst[0] < t)) continue;
                    }
                    set |= 1u << tr.output;
                    // Viterbi_Data_assign (viterbi.c:445-462); shadow stamps are
                    // applied to the transported copy (DESIGN.md "shadows")
                    dst[0] = t;
                    if (start_cb) {
                        for (int l = 1; l < C; ++l) dst[l] = (l <= m.n_shadow_slots) ? src[l] : 0;
                    } else {
                        for (int l = 1; l < C; ++l) dst[l] = src[l];
                    }
                    for (int l = 0; l < m.n_shadow_slots; ++l) {
                        const int kind = m.shadow_start[tr.input][l];
                        if (kind == 1) dst[1 + l] = P.t_start + sj;
                        else if (kind == 2) dst[1 + l] = P.q_start + si;
                    }
                    if (from_start) {
                        if (qid >= 0) dst[qid] = si;
                        if (tid >= 0) dst[tid] = sj;
                    }
                    if (mode == GEN_PATH) P.tb[GEN_TB_CELL(i, j, Q, S) + tr.output] = (uint8_t)k;
                }
                if ((set >> m.end_state) & 1u) {  // viterbi.c:778-791
                    const int32_t *ec = cell + m.end_state * C;
                    const int v = ec[0];
                    // cell_end_func (viterbi.c:792-797) is called by the host binding on these:
                    // Heuristic_Bound_report_end_func / Heuristic_Span_src_report_end_func
                    // (src/bsdp/heuristic.c:139-145,385-410)
                    if (P.end_cells)
                        for (int l = 0; l <= m.n_shadow_slots; ++l)
                            P.end_cells[((size_t)i * (T + 1) + j) * (1 + m.n_shadow_slots) + l] = ec[l];
                    if (v > best || (v == best && (j < best_j || (j == best_j && i < best_i)))) {
                        best = v; best_i = i; best_j = j;
                        best_si = (qid >= 0) ? ec[qid] : 0;
                        best_sj = (tid >= 0) ? ec[tid] : 0;
                    }
                }
            }
            __syncthreads();
        }
        red_score[threadIdx.x] = best; red_i[threadIdx.x] = best_i; red_j[threadIdx.x] = best_j;
        red_si[threadIdx.x] = best_si; red_sj[threadIdx.x] = best_sj;
        __syncthreads();
        for (int off = (int)blockDim.x / 2; off > 0; off >>= 1) {
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
            o.start_i = red_si[0]; o.start_j = red_sj[0];
            o.flags = (red_score[0] == INT32_MIN) ? 1 : 0;
            outs[P.out_index] = o;
        }
        __syncthreads();
    }
}

// REGION results -> PATH lattices restricted to the alignment's bounding box.
__global__ void generic_plan_box_kernel(const GenPair *__restrict__ full, const GenOut *__restrict__ reg,
                                        GenPair *__restrict__ box, int n, int threshold) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    GenPair P = full[p];
    const GenOut o = reg[P.out_index];
    P.blk_dq += o.start_i;
    P.blk_dt += o.start_j;
    P.q_start += o.start_i;
    P.t_start += o.start_j;
    P.Q = o.end_i - o.start_i;
    P.T = o.end_j - o.start_j;
    if (o.score < threshold || o.flags) { P.Q = 0; P.T = 0; }  // nothing to trace
    box[p] = P;
}

__global__ void generic_clear_tb_kernel(const GenPair *__restrict__ pairs, int n, int S) {
    const GenPair P = pairs[blockIdx.y];
    const size_t total = GEN_TB_BYTES(P.Q, P.T, S);
    uint32_t *w = reinterpret_cast<uint32_t *>(P.tb);
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < (total + 3) / 4;
         k += (size_t)gridDim.x * blockDim.x)
        w[k] = 0xFFFFFFFFu;
    (void)n;
}

struct GenJob {
    int32_t pair, result, expect, score_slot;
    int64_t ops_off;
    int32_t ops_cap, reserved;
};

__global__ void generic_traceback_kernel(const GenPair *__restrict__ pairs, const GenOut *__restrict__ outs,
                                         const GenOut *__restrict__ reg_outs, const GenJob *__restrict__ jobs,
                                         int n, const GenTables *__restrict__ tables, int threshold,
                                         c4b_result *__restrict__ results, int32_t *__restrict__ ops) {
    // decode tables of the rank-bit record (GEN_TBS_CHUNK): bit offset / width of every state in a
    // row's record, and rank -> transition id per state; exactly as the host laid them out
    __shared__ short s_bit_off[C4B_MAX_STATES], s_bit_n[C4B_MAX_STATES];
    __shared__ unsigned char s_rank2tr[C4B_MAX_STATES][C4B_MAX_TRANSITIONS + 1];
    __shared__ unsigned char s_in[C4B_MAX_TRANSITIONS], s_aq[C4B_MAX_TRANSITIONS], s_at[C4B_MAX_TRANSITIONS];
    __shared__ int s_row_bits;
    const c4b_model &m = tables->model;
    const int S = m.n_states;
    if (threadIdx.x == 0) {
        int row_bits = 0;
        for (int s = 0; s < S; ++s) {
            int n_in = 0, b = 0;
            for (int k = 0; k < m.n_transitions; ++k)
                if (m.transitions[k].output == s) s_rank2tr[s][++n_in] = (unsigned char)k;
            s_rank2tr[s][0] = 0xFF;
            while ((1 << b) < n_in + 1) ++b;
            s_bit_off[s] = (short)row_bits; s_bit_n[s] = (short)b; row_bits += b;
        }
        s_row_bits = row_bits;
        for (int k = 0; k < m.n_transitions; ++k) {
            s_in[k] = (unsigned char)m.transitions[k].input;
            s_aq[k] = (unsigned char)m.transitions[k].advance_query;
            s_at[k] = (unsigned char)m.transitions[k].advance_target;
        }
    }
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const GenJob J = jobs[g];
    const GenPair P = pairs[J.pair];
    const GenOut o = outs[P.out_index];
    c4b_result res;
    res.score = o.score; res.status = 0; res.reserved = 0; res.n_ops = 0; res.ops_offset = J.ops_off;
    res.query_start = P.q_start; res.target_start = P.t_start;
    res.query_end = P.q_start + o.end_i; res.target_end = P.t_start + o.end_j;
    if (J.expect) {
        const GenOut r = reg_outs[J.score_slot];
        res.score = r.score;
        if (r.score < threshold) { res.status = 1; results[J.result] = res; return; }
        if (r.score != o.score) res.status = 3;  // optimal.c:394-399
    }
    if (o.flags) res.status = 2;
    int32_t *out = ops + 2 * J.ops_off;
    int n_runs = 0, last_t = -1, i = o.end_i, j = o.end_j;
    if (res.status == 0) {
        const int row_bits = s_row_bits;
        auto winner = [&](int ci, int cj, int state) -> int {   // transition id, 0xFF = unset
            if (!P.tb_rows) return P.tb[GEN_TB_CELL(ci, cj, P.Q, S) + state];
            const int nb = s_bit_n[state];
            if (nb == 0) return 0xFF;
            const uint32_t *w = reinterpret_cast<const uint32_t *>(P.tb + GEN_TBS_CHUNK(ci, cj, P.T, P.tb_rows, P.tb_chunk));
            const int pos = (ci % P.tb_rows) * row_bits + s_bit_off[state];
            uint64_t v = w[pos / 32];
            if (pos % 32 + nb > 32) v |= (uint64_t)w[pos / 32 + 1] << 32;
            return s_rank2tr[state][(int)((v >> (pos % 32)) & ((1u << nb) - 1u))];
        };
        auto emit = [&](int tr, int count) -> bool {
            if (tr == last_t) { out[2 * (n_runs - 1) + 1] += count; return true; }
            if (n_runs >= J.ops_cap) return false;
            out[2 * n_runs] = tr; out[2 * n_runs + 1] = count; ++n_runs; last_t = tr;
            return true;
        };
        int tr = winner(i, j, m.end_state);
        while (tr != 0xFF) {
            if (!emit(tr, 1)) { res.status = 4; break; }
            const int aq = s_aq[tr], at = s_at[tr], in = s_in[tr];
            i -= aq;
            j -= at;
            if (in == m.start_state) break;
            if (i < 0 || j < 0) { res.status = 4; break; }
            if (m.transitions[tr].output == in && aq + at > 0) {
                // a self-loop (an intron, a gap run): the walk is a chain of dependent loads, one per
                // cell, but while the loop goes on the cells to look at are known in advance -- fetch
                // eight at a time and consume them in order
                for (;;) {
                    int nxt[8], kk = 0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int ci = i - u * aq, cj = j - u * at;
                        nxt[u] = (ci >= 0 && cj >= 0) ? winner(ci, cj, in) : 0xFF;
                    }
                    while (kk < 8 && nxt[kk] == tr) ++kk;
                    if (kk) {
                        if (!emit(tr, kk)) { res.status = 4; break; }
                        i -= kk * aq;
                        j -= kk * at;
                    }
                    if (kk < 8) {
                        // the loop ended: cell (i, j) holds another winner (or the lattice edge)
                        tr = (i >= 0 && j >= 0) ? nxt[kk] : 0xFF;
                        if (i < 0 || j < 0) res.status = 4;
                        break;
                    }
                }
                if (res.status) break;
                continue;   // `tr` is the winner of state `in` at the current cell
            }
            tr = winner(i, j, in);
        }
        for (int a = 0, b = n_runs - 1; a < b; ++a, --b) {
            const int t0 = out[2 * a], l0 = out[2 * a + 1];
            out[2 * a] = out[2 * b]; out[2 * a + 1] = out[2 * b + 1];
            out[2 * b] = t0; out[2 * b + 1] = l0;
        }
        res.query_start = P.q_start + max(i, 0);
        res.target_start = P.t_start + max(j, 0);
    }
    res.n_ops = n_runs;
    results[J.result] = res;
}

// ---- windowed traceback of lattices whose whole PATH record would not fit (JIT_SYS_WIN) ------
// Cursors start at the END cell pass 1 found (checks of generic_traceback_kernel: optimal.c:394-411).
__global__ void generic_window_walk_init_kernel(const GenPair *__restrict__ pairs, const GenOut *__restrict__ outs,
                                                const GenOut *__restrict__ reg_outs, const GenJob *__restrict__ jobs,
                                                const int32_t *__restrict__ big, int n, const GenTables *__restrict__ tables,
                                                int threshold, GenWalk *__restrict__ walk,
                                                c4b_result *__restrict__ results) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const GenJob J = jobs[big[g]];
    const GenPair P = pairs[g];
    const GenOut o = outs[P.out_index];
    c4b_result res;
    res.score = o.score; res.status = 0; res.reserved = 0; res.n_ops = 0; res.ops_offset = J.ops_off;
    res.query_start = P.q_start; res.target_start = P.t_start;
    res.query_end = P.q_start + o.end_i; res.target_end = P.t_start + o.end_j;
    if (J.expect) {
        const GenOut r = reg_outs[J.score_slot];
        res.score = r.score;
        if (r.score < threshold) res.status = 1;
        else if (r.score != o.score) res.status = 3;
    }
    if (res.status == 0 && o.flags) res.status = 2;
    GenWalk W;
    W.i = o.end_i; W.j = o.end_j; W.state = tables->model.end_state;
    W.n_runs = 0; W.last_t = -1; W.status = res.status; W.done = res.status != 0; W.reserved = 0;
    walk[g] = W;
    results[J.result] = res;   // (final for rejected lattices; the walk overwrites it when it finishes)
}

// Viterbi_Data_create_Alignment (viterbi.c:342-392) inside the window that was just refilled: one
// thread per lattice walks END -> START until the cursor leaves the window on its left (the next
// round refills the window it is then in) or reaches START.
__global__ void generic_window_walk_kernel(const GenPair *__restrict__ pairs, const GenOut *__restrict__ outs,
                                           const GenJob *__restrict__ jobs, const int32_t *__restrict__ big,
                                           const GenWin *__restrict__ wins, int n,
                                           const GenTables *__restrict__ tables, GenWalk *__restrict__ walk,
                                           c4b_result *__restrict__ results, int32_t *__restrict__ ops) {
    __shared__ short s_bit_off[C4B_MAX_STATES], s_bit_n[C4B_MAX_STATES];
    __shared__ unsigned char s_rank2tr[C4B_MAX_STATES][C4B_MAX_TRANSITIONS + 1];
    __shared__ int s_row_bits;
    const c4b_model &m = tables->model;
    const int S = m.n_states;
    if (threadIdx.x == 0) {   // the rank-bit record's decode tables, as in generic_traceback_kernel
        int row_bits = 0;
        for (int s = 0; s < S; ++s) {
            int n_in = 0, b = 0;
            for (int k = 0; k < m.n_transitions; ++k)
                if (m.transitions[k].output == s) s_rank2tr[s][++n_in] = (unsigned char)k;
            s_rank2tr[s][0] = 0xFF;
            while ((1 << b) < n_in + 1) ++b;
            s_bit_off[s] = (short)row_bits; s_bit_n[s] = (short)b; row_bits += b;
        }
        s_row_bits = row_bits;
    }
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    GenWalk W = walk[g];
    if (W.done) return;
    const GenJob J = jobs[big[g]];
    const GenPair P = pairs[g];
    const GenWin win = wins[g];
    int32_t *out = ops + 2 * J.ops_off;
    int i = W.i, j = W.j, state = W.state, n_runs = W.n_runs, last_t = W.last_t, status = 0;
    bool finished = false;
    const int row_bits = s_row_bits;
    auto winner = [&](int ci, int cj, int st) -> int {   // records of the window: a lattice of wcols columns
        const int nb = s_bit_n[st];
        if (nb == 0) return 0xFF;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(
            P.tb + GEN_TBS_CHUNK(ci, cj - win.c0, win.wcols - 1, P.tb_rows, P.tb_chunk));
        const int pos = (ci % P.tb_rows) * row_bits + s_bit_off[st];
        uint64_t v = w[pos / 32];
        if (pos % 32 + nb > 32) v |= (uint64_t)w[pos / 32 + 1] << 32;
        return s_rank2tr[st][(int)((v >> (pos % 32)) & ((1u << nb) - 1u))];
    };
    while (j >= win.c0) {
        const int tr = winner(i, j, state);
        if (tr == 0xFF) { finished = true; break; }
        if (tr == last_t) out[2 * (n_runs - 1) + 1] += 1;
        else if (n_runs < J.ops_cap) { out[2 * n_runs] = tr; out[2 * n_runs + 1] = 1; ++n_runs; last_t = tr; }
        else { status = 4; break; }
        const c4b_transition &t = m.transitions[tr];
        i -= t.advance_query;
        j -= t.advance_target;
        if (t.input == m.start_state) { finished = true; break; }
        if (i < 0 || j < 0) { status = 4; break; }
        state = t.input;
    }
    if (finished || status) {
        const GenOut o = outs[P.out_index];
        c4b_result res = results[J.result];   // score / end as the init kernel left them
        res.status = status;
        if (!status) {
            for (int a = 0, b = n_runs - 1; a < b; ++a, --b) {
                const int t0 = out[2 * a], l0 = out[2 * a + 1];
                out[2 * a] = out[2 * b]; out[2 * a + 1] = out[2 * b + 1];
                out[2 * b] = t0; out[2 * b + 1] = l0;
            }
            res.query_start = P.q_start + max(i, 0);
            res.target_start = P.t_start + max(j, 0);
        }
        res.query_end = P.q_start + o.end_i; res.target_end = P.t_start + o.end_j;
        res.n_ops = n_runs;
        results[J.result] = res;
        W.done = 1;
        W.status = status;
    }
    W.i = i; W.j = j; W.state = state; W.n_runs = n_runs; W.last_t = last_t;
    walk[g] = W;
}

__global__ void generic_score_results_kernel(const GenPair *__restrict__ pairs, const GenOut *__restrict__ outs,
                                             int n, c4b_result *__restrict__ results) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const GenPair P = pairs[p];
    const GenOut o = outs[P.out_index];
    c4b_result r;
    r.score = o.score;
    r.query_start = P.q_start; r.target_start = P.t_start;
    r.query_end = P.q_start + o.end_i; r.target_end = P.t_start + o.end_j;
    r.n_ops = 0; r.ops_offset = 0; r.status = o.flags ? 2 : 0; r.reserved = 0;
    results[P.out_index] = r;
}

// ---- host side ---------------------------------------------------------------
struct GenericBatch;
// device copies of caller-declared stable host buffers: (host address, bytes) -> device address
struct ResidentBuffers {
    std::map<std::pair<const void *, size_t>, void *> map;
    size_t bytes = 0;
};
// per-lattice START / END cell tables that already live on the device (batched span fills):
// host arrays of n device pointers, either may be null
struct GenDevTables {
    const int32_t *const *start = nullptr;
    int32_t *const *end = nullptr;
};
int generic_batch_create(cudaStream_t stream, int64_t *launch_counter, const c4b_model *model,
                         const c4b_scoring *scoring, int n, const c4b_pair *pairs, bool want_path,
                         GenericBatch **out, const int32_t *start_cells = nullptr, bool end_cells = false,
                         int sm_count = 0, ResidentBuffers *resident = nullptr, const GenDevTables *dev = nullptr);
int generic_batch_run(GenericBatch *g, c4b_score threshold);
int generic_batch_fetch(GenericBatch *g, c4b_result *results, int32_t *ops, int64_t ops_capacity);
int64_t generic_batch_cells(const GenericBatch *g);
const void *generic_batch_device_results(const GenericBatch *g);
double generic_batch_fill_ms(GenericBatch *g);
void generic_batch_destroy(GenericBatch *g);

}  // namespace c4b
