/* c4_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see c4_oracle.h).
 *
 * Restates, in plain C over the flat tables of include/c4b200.h, what the
 * reference computes on this path.  Each function cites the reference code it
 * follows (paths relative to /root/reference).  Nothing here is used by the
 * product; the CUDA library is checked AGAINST this.
 */
#include "c4_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LOW C4B_IMPOSSIBLY_LOW_SCORE
#define HIGH C4B_IMPOSSIBLY_HIGH_SCORE

/* ---- Layout (src/c4/layout.c:21-88): a state is "active" at a lattice point
 * if the point is inside the lattice and, for START / END, inside its scope. */
static int state_active(const c4b_model *m, int state, int qp, int tp, int ql, int tl) {
    if (qp < 0 || tp < 0 || qp > ql || tp > tl) return 0;
    if (state == m->start_state) {
        switch (m->start_scope) {
        case C4B_SCOPE_ANYWHERE: break;
        case C4B_SCOPE_EDGE: if (qp != 0 && tp != 0) return 0; break;
        case C4B_SCOPE_QUERY: if (qp != 0) return 0; break;
        case C4B_SCOPE_TARGET: if (tp != 0) return 0; break;
        case C4B_SCOPE_CORNER: if (qp != 0 || tp != 0) return 0; break;
        default: return 0;
        }
    }
    if (state == m->end_state) {
        switch (m->end_scope) {
        case C4B_SCOPE_ANYWHERE: break;
        case C4B_SCOPE_EDGE: if (qp != ql && tp != tl) return 0; break;
        case C4B_SCOPE_QUERY: if (qp != ql) return 0; break;
        case C4B_SCOPE_TARGET: if (tp != tl) return 0; break;
        case C4B_SCOPE_CORNER: if (qp != ql || tp != tl) return 0; break;
        default: return 0;
        }
    }
    return 1;
}

/* src/c4/layout.c:122-154 (the third check is disabled in the reference). */
int c4o_transition_is_valid(const c4b_model *m, int k, int i, int j, int ql, int tl) {
    const c4b_transition *t = &m->transitions[k];
    return state_active(m, t->input, i - t->advance_query, j - t->advance_target, ql, tl) &&
           state_active(m, t->output, i, j, ql, tl);
}

/* ---- per-cell scoring (C4_Calc_score, src/c4/c4.c:1700-1711 + callbacks) -- */
static int submat(const int32_t *matrix, const uint8_t *index, int a, int b) {
    /* Submat_lookup, src/sequence/submat.h:54-56 */
    return matrix[index[a & 255] * C4B_SUBMAT_N + index[b & 255]];
}
static int translate(const c4b_scoring *s, int a, int b, int c) {
    /* Translate_base, src/sequence/translate.h:76-79 */
    return s->codon_aa[s->nt2d[a & 255] | (s->nt2d[b & 255] << 4) | (s->nt2d[c & 255] << 8)];
}

c4b_score c4o_calc_score(const c4b_model *m, const c4b_scoring *s, const c4b_pair *p,
                         int calc_id, int qp, int tp, const c4b_score *src) {
    const c4b_calc *c;
    const uint8_t *q = p->query, *t = p->target;
    int slot, len;
    if (calc_id < 0) return 0; /* NULL calc scores zero */
    c = &m->calcs[calc_id];
    switch (c->kind) {
    case C4B_CALC_CONST: return c->param[0];
    case C4B_CALC_MATCH_DNA: /* match.c:271-285 (no CDS annotation support) */
        return submat(s->dna_matrix, s->dna_index, q[qp], t[tp]);
    case C4B_CALC_MATCH_PROTEIN: /* match.c:287-295 */
        return submat(s->protein_matrix, s->protein_index, q[qp], t[tp]);
    case C4B_CALC_MATCH_1_3: /* match.c:332-355 */
        return submat(s->protein_matrix, s->protein_index, q[qp],
                      translate(s, t[tp], t[tp + 1], t[tp + 2]));
    case C4B_CALC_MATCH_3_1:
        return submat(s->protein_matrix, s->protein_index,
                      translate(s, q[qp], q[qp + 1], q[qp + 2]), t[tp]);
    case C4B_CALC_MATCH_3_3: /* match.c:508-530 */
        return submat(s->protein_matrix, s->protein_index,
                      translate(s, q[qp], q[qp + 1], q[qp + 2]),
                      translate(s, t[tp], t[tp + 1], t[tp + 2]));
    case C4B_CALC_SPLICE_PRE: /* intron.c:138-161 with is_pre */
        return c->param[0] + p->splice[c->param[1]][tp];
    case C4B_CALC_SPLICE_POST: /* intron.c:151-159 */
        slot = src[1 + c->param[2]];
        len = tp - slot + 2;
        if (len < s->min_intron || len > s->max_intron) return LOW;
        return p->splice[c->param[1]][tp];
    case C4B_CALC_PHASE1_POST: /* phase.c:145-163,174-186 */
        slot = src[1 + c->param[2]];
        if (slot < 1) return LOW;
        return submat(s->protein_matrix, s->protein_index, q[qp],
                      translate(s, t[slot - 1], t[tp], t[tp + 1]));
    case C4B_CALC_PHASE2_POST:
        slot = src[1 + c->param[2]];
        if (slot < 2) return LOW;
        return submat(s->protein_matrix, s->protein_index, q[qp],
                      translate(s, t[slot - 2], t[slot - 1], t[tp]));
    default: return LOW;
    }
}

/* ---- SubOpt_Index (src/c4/subopt.c:250-374): exact-set lookup ----------- */
static int is_blocked(const c4b_pair *p, int i, int j) {
    int lo = 0, hi = p->n_blocked;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        int tj = p->blocked_target_pos[mid], qi = p->blocked_query_pos[mid];
        if (tj < j || (tj == j && qi < i)) lo = mid + 1;
        else hi = mid;
    }
    return lo < p->n_blocked && p->blocked_target_pos[lo] == j &&
           p->blocked_query_pos[lo] == i;
}

/* ---- Alignment_add (src/c4/alignment.c:75-100): RLE append ------------- */
static int ops_add(int32_t *ops, int64_t cap, int32_t *n_ops, int transition) {
    if (*n_ops && ops[2 * (*n_ops - 1)] == transition) {
        ops[2 * (*n_ops - 1) + 1]++;
        return 0;
    }
    if (*n_ops >= cap) return -1;
    ops[2 * *n_ops] = transition;
    ops[2 * *n_ops + 1] = 1;
    (*n_ops)++;
    return 0;
}

/* ---- the fill: Viterbi_interpreted, src/c4/viterbi.c:655-837 ----------- */
int c4o_viterbi(const c4b_model *m, const c4b_scoring *s, const c4b_pair *p, int mode,
                c4b_result *res, int32_t *ops, int64_t ops_capacity) {
    return c4o_viterbi_cells(m, s, p, mode, NULL, NULL, res, ops, ops_capacity);
}

/* ... with the START / END cell callbacks of BSDP derived models as tables
 * (cell_start_func viterbi.c:727-741, cell_end_func :792-797; see c4b_viterbi_calculate_cells) */
int c4o_viterbi_cells(const c4b_model *m, const c4b_scoring *s, const c4b_pair *p, int mode,
                      const c4b_score *start_cells, c4b_score *end_cells,
                      c4b_result *res, int32_t *ops, int64_t ops_capacity) {
    const int S = m->n_states, Tn = m->n_transitions;
    c4b_score dummy_start[1 + C4B_MAX_SHADOW_SLOTS + 2];
    const int ql = p->query_length, tl = p->target_length;
    const int mta = m->max_target_advance;
    int C = 1 + m->n_shadow_slots; /* Viterbi_get_cell_size, viterbi.c:42-56 */
    int qid = -1, tid = -1;
    int i, j, k, l, is_set[C4B_MAX_STATES];
    c4b_score *rows, score = LOW;
    int8_t *tb = NULL;
    int end_is_set = 0, end_i = 0, end_j = 0, start_i = 0, start_j = 0;
    size_t row_stride, cell_stride;

    if (S > C4B_MAX_STATES || Tn > C4B_MAX_TRANSITIONS) return -2;
    if (mode == C4O_FIND_REGION && m->start_scope != C4B_SCOPE_CORNER) {
        /* Viterbi_Row_create, viterbi.c:163-170 */
        if (m->start_scope != C4B_SCOPE_QUERY) qid = C++;
        if (m->start_scope != C4B_SCOPE_TARGET) tid = C++;
    }
    cell_stride = (size_t)C;
    row_stride = (size_t)(ql + 1) * S * cell_stride;
    rows = (c4b_score *)malloc(sizeof(c4b_score) * row_stride * (size_t)(mta + 1));
    if (!rows) return -3;
    for (l = 0; l < (int)(row_stride * (size_t)(mta + 1)); l++) rows[l] = 0;
    for (l = 0; l < (mta + 1) * (ql + 1) * S; l++) rows[(size_t)l * cell_stride] = LOW;
    if (mode == C4O_FIND_PATH) {
        /* Viterbi_traceback_memory_create, viterbi.c:220-227 (ids, not pointers) */
        size_t n = (size_t)(ql + 1) * (size_t)(tl + 1) * (size_t)S;
        tb = (int8_t *)malloc(n);
        if (!tb) { free(rows); return -3; }
        memset(tb, -1, n);
    }
#define ROW(jj) (rows + (size_t)((jj) % (mta + 1)) * row_stride)
#define CELL(jj, ii, st) (ROW(jj) + ((size_t)(ii) * S + (st)) * cell_stride)
    for (j = 0; j <= tl; j++) {
        for (i = 0; i <= ql; i++) {
            for (k = 0; k < S; k++) { /* viterbi.c:691-694 */
                is_set[k] = 0;
                CELL(j, i, k)[0] = LOW;
            }
            for (k = 0; k < Tn; k++) {
                const c4b_transition *tr = &m->transitions[k];
                c4b_score t, *src, *dst;
                int sq = i - tr->advance_query, st = j - tr->advance_target;
                if (!c4o_transition_is_valid(m, k, i, j, ql, tl)) continue; /* :697-700 */
                if (tr->label == C4B_LABEL_MATCH && p->n_blocked && is_blocked(p, i, j))
                    continue; /* :701-704 */
                dst = CELL(j, i, tr->output);
                if (tr->input == m->start_state && start_cells) {
                    /* cell_start_func's cell is copied and stands in for the source (:727-741) */
                    for (l = 0; l < C; l++)
                        dummy_start[l] = (l <= m->n_shadow_slots)
                            ? start_cells[((size_t)sq * (tl + 1) + st) * (1 + m->n_shadow_slots) + l] : 0;
                    src = dummy_start;
                    t = src[0];
                } else if (tr->input == m->start_state) {
                    /* t = 0 without a cell_start_func (:720-745); the START
                     * state's own row cell supplies the shadow slots. */
                    src = CELL(st, sq, tr->input);
                    t = 0;
                } else {
                    src = CELL(st, sq, tr->input);
                    t = src[0];
                }
                /* shadow end + calc at SOURCE coordinates (:747-752) */
                t += c4o_calc_score(m, s, p, tr->calc, p->query_start + sq,
                                    p->target_start + st, src);
                if (tr->calc >= 0) { /* :754-765 */
                    int prot = m->calcs[tr->calc].protect;
                    if ((prot & C4B_PROTECT_UNDERFLOW) && t < LOW) t = LOW;
                    if ((prot & C4B_PROTECT_OVERFLOW) && t > HIGH) t = HIGH;
                }
                if (is_set[tr->output] && !(dst[0] < t)) continue; /* :766-775 */
                is_set[tr->output] = 1;
                /* Viterbi_Data_assign, viterbi.c:445-462 */
                dst[0] = t;
                if (tr->input == m->start_state) { /* Viterbi_Row_shadow_start :403-411 */
                    if (qid >= 0) src[qid] = sq;
                    if (tid >= 0) src[tid] = st;
                }
                for (l = 0; l < m->n_shadow_slots; l++) { /* :413-422 */
                    int kind = m->shadow_start[tr->input][l];
                    if (kind == 1) src[1 + l] = p->target_start + st;
                    else if (kind == 2) src[1 + l] = p->query_start + sq;
                }
                for (l = 1; l < C; l++) dst[l] = src[l]; /* shadow transport :456-457 */
                if (tb) tb[((size_t)i * (tl + 1) + j) * S + tr->output] = (int8_t)k;
            }
            if (is_set[m->end_state]) { /* :778-791, Viterbi_Data_register_end :464-478 */
                c4b_score *cell = CELL(j, i, m->end_state);
                if (end_cells) /* what cell_end_func is handed (:792-797) */
                    for (l = 0; l <= m->n_shadow_slots; l++)
                        end_cells[((size_t)i * (tl + 1) + j) * (1 + m->n_shadow_slots) + l] = cell[l];
                if (!end_is_set || score < cell[0]) {
                    score = cell[0];
                    end_is_set = 1;
                    end_i = i;
                    end_j = j;
                    if (qid >= 0) start_i = cell[qid];
                    if (tid >= 0) start_j = cell[tid];
                }
            }
        }
    }
    res->score = score;
    res->status = end_is_set ? 0 : 2;
    res->n_ops = 0;
    res->ops_offset = 0;
    res->reserved = 0;
    res->query_start = p->query_start;
    res->target_start = p->target_start;
    res->query_end = p->query_start + end_i;
    res->target_end = p->target_start + end_j;
    if (mode == C4O_FIND_REGION) { /* Viterbi_Data_finalise, viterbi.c:633-653 */
        if (qid >= 0) res->query_start = p->query_start + start_i;
        if (tid >= 0) res->target_start = p->target_start + start_j;
    }
    if (mode == C4O_FIND_PATH && end_is_set) {
        /* Viterbi_Data_create_Alignment, viterbi.c:342-392 */
        int n = 0, cap = 1024, *path = (int *)malloc(sizeof(int) * 1024);
        int32_t n_ops = 0;
        int tr = tb[((size_t)end_i * (tl + 1) + end_j) * S + m->end_state];
        i = end_i;
        j = end_j;
        while (tr >= 0) {
            if (n == cap) path = (int *)realloc(path, sizeof(int) * (size_t)(cap *= 2));
            path[n++] = tr;
            i -= m->transitions[tr].advance_query;
            j -= m->transitions[tr].advance_target;
            if (m->transitions[tr].input == m->start_state) break;
            tr = tb[((size_t)i * (tl + 1) + j) * S + m->transitions[tr].input];
        }
        res->query_start = p->query_start + i;
        res->target_start = p->target_start + j;
        for (k = n - 1; k >= 0; k--)
            if (ops_add(ops, ops_capacity, &n_ops, path[k])) {
                free(path); free(tb); free(rows);
                return -4;
            }
        res->n_ops = n_ops;
        free(path);
    }
#undef ROW
#undef CELL
    free(tb);
    free(rows);
    return 0;
}

/* Optimal_find_path, src/c4/optimal.c:368-413 */
int c4o_find_path(const c4b_model *m, const c4b_scoring *s, const c4b_pair *p,
                  c4b_score threshold, int64_t region_threshold_cells, c4b_result *res,
                  int32_t *ops, int64_t ops_capacity) {
    int64_t cells = (int64_t)(p->query_length + 1) * (p->target_length + 1);
    int global = (m->start_scope == C4B_SCOPE_CORNER && m->end_scope == C4B_SCOPE_CORNER);
    int rc;
    if (cells > region_threshold_cells && !global) {
        c4b_result reg;
        c4b_pair sub = *p;
        rc = c4o_viterbi(m, s, p, C4O_FIND_REGION, &reg, NULL, 0);
        if (rc) return rc;
        if (reg.score < threshold) {
            *res = reg;
            res->status = 1;
            return 0;
        }
        sub.query_start = reg.query_start;
        sub.target_start = reg.target_start;
        sub.query_length = reg.query_end - reg.query_start;
        sub.target_length = reg.target_end - reg.target_start;
        /* NB: blocked cells are region-relative; callers of the oracle pass
         * n_blocked = 0 on this branch or pre-shift them. */
        rc = c4o_viterbi(m, s, &sub, C4O_FIND_PATH, res, ops, ops_capacity);
        if (rc) return rc;
        if (res->score != reg.score) return -5; /* optimal.c:394-399 */
    } else {
        rc = c4o_viterbi(m, s, p, C4O_FIND_PATH, res, ops, ops_capacity);
        if (rc) return rc;
    }
    if (res->score < threshold) res->status = 1;
    return 0;
}

/* Alignment_has_valid_alignment, src/c4/alignment.c:3240-3372 */
c4b_score c4o_rescore_path(const c4b_model *m, const c4b_scoring *s, const c4b_pair *p,
                           const c4b_result *res, const int32_t *ops) {
    c4b_score cell[1 + C4B_MAX_SHADOW_SLOTS] = {0}, score = 0;
    int qp = res->query_start, tp = res->target_start, i, r, l;
    for (i = 0; i < res->n_ops; i++) {
        const c4b_transition *tr = &m->transitions[ops[2 * i]];
        for (r = 0; r < ops[2 * i + 1]; r++) {
            for (l = 0; l < m->n_shadow_slots; l++) {
                int kind = m->shadow_start[tr->input][l];
                if (kind == 1) cell[1 + l] = tp;
                else if (kind == 2) cell[1 + l] = qp;
            }
            score += c4o_calc_score(m, s, p, tr->calc, qp, tp, cell);
            qp += tr->advance_query;
            tp += tr->advance_target;
        }
    }
    return score;
}

/* ---- report strings ----------------------------------------------------- */
typedef struct {
    char *buf;
    int len, cap, overflow;
} sbuf;
static void sb_put(sbuf *b, const char *fmt, int a0, int a1, int a2) {
    int room = b->cap - b->len, n;
    if (b->overflow) return;
    n = snprintf(b->buf + b->len, (size_t)(room > 0 ? room : 0), fmt, a0, a1, a2);
    if (n < 0 || n >= room) b->overflow = 1;
    else b->len += n;
}

/* Alignment_print_cigar_block, alignment.c:1641-1681 */
int c4o_format_cigar(const c4b_model *m, const int32_t *ops, int n_ops, char *buf, int buflen) {
    sbuf b = {buf, 0, buflen, 0};
    int i, first = 1, type = 0, move = 0;
    if (buflen > 0) buf[0] = '\0';
    for (i = 0; i < n_ops; i++) {
        const c4b_transition *tr = &m->transitions[ops[2 * i]];
        int len = ops[2 * i + 1], ty, mv;
        if (!tr->advance_query) { ty = 'D'; mv = tr->advance_target * len; }
        else if (!tr->advance_target) { ty = 'I'; mv = tr->advance_query * len; }
        else { ty = 'M'; mv = (tr->advance_query > tr->advance_target ? tr->advance_query : tr->advance_target) * len; }
        if (i == 0) { type = ty; move = mv; continue; }
        if (ty == type) { move += mv; continue; }
        if (move) { sb_put(&b, first ? "%c %d" : " %c %d", type, move, 0); }
        /* the reference sets the separator after the first type CHANGE, even if
         * nothing was printed for a zero-length head (alignment.c:1667-1673) */
        first = 0;
        type = ty;
        move = mv;
    }
    if (n_ops && move) sb_put(&b, first ? "%c %d" : " %c %d", type, move, 0);
    return b.overflow ? -1 : b.len;
}

/* Alignment_print_vulgar_block, alignment.c:1683-1769 */
int c4o_format_vulgar(const c4b_model *m, const int32_t *ops, int n_ops, char *buf, int buflen) {
    static const char label_char[] = {0, 'M', 'G', 'N', '5', '3', 'I', 'S', 'F'};
    sbuf b = {buf, 0, buflen, 0};
    int i, first = 1, label, aq, at, is_codon = 0;
    if (buflen > 0) buf[0] = '\0';
    if (!n_ops) return 0;
    label = m->transitions[ops[0]].label;
    aq = m->transitions[ops[0]].advance_query * ops[1];
    at = m->transitions[ops[0]].advance_target * ops[1];
    for (i = 1; i < n_ops; i++) {
        const c4b_transition *tr = &m->transitions[ops[2 * i]];
        int len = ops[2 * i + 1];
        int codon = (tr->advance_query == 3 && tr->advance_target == 3);
        if (tr->label == label && (aq || !tr->advance_query) && (at || !tr->advance_target) &&
            is_codon == codon) {
            aq += tr->advance_query * len;
            at += tr->advance_target * len;
            continue;
        }
        if (label != C4B_LABEL_NONE) {
            int ch = (label == C4B_LABEL_MATCH && is_codon) ? 'C' : label_char[label];
            sb_put(&b, first ? "%c %d %d" : " %c %d %d", ch, aq, at);
            first = 0;
        }
        label = tr->label;
        is_codon = codon;
        aq = tr->advance_query * len;
        at = tr->advance_target * len;
    }
    /* the trailing block is NOT flushed by the reference: the final
     * "match to end" (label NONE) transition is what pushes the last real
     * block out (alignment.c:1696-1766). */
    return b.overflow ? -1 : b.len;
}

/* ======================================================================
 * HSP seeding / extension -- restatement of src/comparison/hspset.c
 * (TEST INFRASTRUCTURE: the checker of c4b_hsp_extend_batch)
 * ====================================================================== */
typedef struct {
    const c4b_scoring *s;
    const c4b_hsp_param *p;
    const uint8_t *q, *t, *qm, *tm;
    int ql, tl, qadv, tadv;
} hsp_ctx;

/* HSP_get_score -> Match.score_func (match.c:271-295,332-355) */
static int hsp_score(const hsp_ctx *c, int qp, int tp) {
    switch (c->p->match_kind) {
    case C4B_CALC_MATCH_DNA: return submat(c->s->dna_matrix, c->s->dna_index, c->q[qp], c->t[tp]);
    case C4B_CALC_MATCH_PROTEIN: return submat(c->s->protein_matrix, c->s->protein_index, c->q[qp], c->t[tp]);
    case C4B_CALC_MATCH_1_3:
        return submat(c->s->protein_matrix, c->s->protein_index, c->q[qp],
                      translate(c->s, c->t[tp], c->t[tp + 1], c->t[tp + 2]));
    case C4B_CALC_MATCH_3_1:
        return submat(c->s->protein_matrix, c->s->protein_index,
                      translate(c->s, c->q[qp], c->q[qp + 1], c->q[qp + 2]), c->t[tp]);
    default: /* 3:3, match.c:508-530 */
        return submat(c->s->protein_matrix, c->s->protein_index,
                      translate(c->s, c->q[qp], c->q[qp + 1], c->q[qp + 2]),
                      translate(c->s, c->t[tp], c->t[tp + 1], c->t[tp + 2]));
    }
}
/* Match_1_mask_func / Match_3_mask_func (match.c:178-183,212-220) */
static int hsp_qmasked(const hsp_ctx *c, int qp) {
    int k;
    if (!c->qm) return 0;
    for (k = 0; k < c->qadv; ++k)
        if (c->qm[qp + k]) return 1;
    return 0;
}
static int hsp_tmasked(const hsp_ctx *c, int tp) {
    int k;
    if (!c->tm) return 0;
    for (k = 0; k < c->tadv; ++k)
        if (c->tm[tp + k]) return 1;
    return 0;
}

/* HSP_extend, hspset.c:747-812 */
static void hsp_extend(const hsp_ctx *c, c4b_hsp *h, int forbid_masked) {
    int score, maxscore, qp, tp, extend, maxext;
    maxscore = score = h->score;
    qp = h->query_start - c->qadv;
    tp = h->target_start - c->tadv;
    for (extend = 1, maxext = 0; qp >= 0 && tp >= 0; extend++) {
        if (forbid_masked && (hsp_qmasked(c, qp) || hsp_tmasked(c, tp))) break;
        score += hsp_score(c, qp, tp);
        if (maxscore <= score) {
            maxscore = score;
            maxext = extend;
        } else {
            if (score < 0) break;
            if (maxscore - score >= c->p->dropoff) break;
        }
        qp -= c->qadv;
        tp -= c->tadv;
    }
    qp = h->query_start + h->length * c->qadv;
    tp = h->target_start + h->length * c->tadv;
    h->query_start -= maxext * c->qadv;
    h->target_start -= maxext * c->tadv;
    h->length += maxext;
    score = maxscore;
    for (extend = 1, maxext = 0; qp + c->qadv <= c->ql && tp + c->tadv <= c->tl; extend++) {
        if (forbid_masked && (hsp_qmasked(c, qp) || hsp_tmasked(c, tp))) break;
        score += hsp_score(c, qp, tp);
        if (maxscore <= score) {
            maxscore = score;
            maxext = extend;
        } else {
            if (score < 0) break;
            if (maxscore - score >= c->p->dropoff) break;
        }
        qp += c->qadv;
        tp += c->tadv;
    }
    h->score = maxscore;
    h->length += maxext;
}

/* The per-seed part of HSPset_seed_hsp (hspset.c:974-996): trim, init, extend x2,
 * threshold, cobs.  Same outputs as the device entry point. */
void c4o_hsp_extend_one(const c4b_scoring *s, const c4b_hsp_param *p, const uint8_t *q, int ql,
                        const uint8_t *qm, const uint8_t *t, int tl, const uint8_t *tm,
                        c4b_hsp_seed seed, c4b_hsp *h) {
    hsp_ctx c;
    int i, qp, tp, score;
    c.s = s; c.p = p; c.q = q; c.t = t; c.qm = qm; c.tm = tm; c.ql = ql; c.tl = tl;
    c.qadv = (p->match_kind == C4B_CALC_MATCH_3_1 || p->match_kind == C4B_CALC_MATCH_3_3) ? 3 : 1;
    c.tadv = (p->match_kind == C4B_CALC_MATCH_1_3 || p->match_kind == C4B_CALC_MATCH_3_3) ? 3 : 1;
    memset(h, 0, sizeof(*h));
    h->query_start = seed.query_start;
    h->target_start = seed.target_start;
    h->length = p->seedlen;
    /* HSP_trim_ends, hspset.c:850-878 */
    for (i = 0; i < h->length; i++) {
        if (hsp_score(&c, h->query_start, h->target_start) > 0) break;
        h->query_start += c.qadv;
        h->target_start += c.tadv;
    }
    h->length -= i;
    qp = h->query_start + h->length * c.qadv - c.qadv;
    tp = h->target_start + h->length * c.tadv - c.tadv;
    while (h->length > 0) {
        if (hsp_score(&c, qp, tp) > 0) break;
        h->length--;
        qp -= c.qadv;
        tp -= c.tadv;
    }
    /* HSP_init, hspset.c:725-745 */
    qp = h->query_start;
    tp = h->target_start;
    for (i = 0; i < h->length; i++) {
        h->score += hsp_score(&c, qp, tp);
        qp += c.qadv;
        tp += c.tadv;
    }
    if (h->score < 0) {
        h->status = 1; /* g_error("Initial HSP score [%d] less than zero") */
        h->target_end = h->target_start + h->length * c.tadv;
        return;
    }
    /* mask_func is set for every Match_Strand (match.c:670,679): the masked pass always runs */
    hsp_extend(&c, h, 1);
    if (h->score >= p->threshold) hsp_extend(&c, h, 0);
    h->target_end = h->target_start + h->length * c.tadv;
    h->stored = h->score >= p->threshold; /* HSP_store, hspset.c:893-894 */
    if (h->stored) {
        /* HSP_find_cobs, hspset.c:426-441 */
        qp = h->query_start;
        tp = h->target_start;
        score = 0;
        for (i = 0; i < h->length; i++) {
            score += hsp_score(&c, qp, tp);
            if (score >= (h->score >> 1)) break;
            qp += c.qadv;
            tp += c.tadv;
        }
        h->cobs = i;
    }
}

/* HSPset_seed_hsp over a seed list with the diagonal horizon (hspset.c:933-997,
 * seed_repeat == 1, filter_threshold == 0), then HSPset_finalise: the stored HSPs in
 * seed order.  `ext` are the per-seed results (from c4o_hsp_extend_one or from the
 * device); returns the number of HSPs written to out. */
int c4o_hspset_replay(const c4b_hsp_param *p, int ql, int n_seeds, const c4b_hsp_seed *seeds,
                      const c4b_hsp *ext, c4b_hsp *out) {
    const int qadv = (p->match_kind == C4B_CALC_MATCH_3_1 || p->match_kind == C4B_CALC_MATCH_3_3) ? 3 : 1;
    const int tadv = (p->match_kind == C4B_CALC_MATCH_1_3 || p->match_kind == C4B_CALC_MATCH_3_3) ? 3 : 1;
    int *horizon = (int *)calloc((size_t)ql * qadv * tadv, sizeof(int));
    int k, n = 0;
    for (k = 0; k < n_seeds; ++k) {
        const int qs = seeds[k].query_start, ts = seeds[k].target_start;
        const int diag = ts * qadv - qs * tadv;
        const int section = ((diag + ql) % ql + ql) % ql; /* (diag_pos + query->len) % query->len, C remainder kept non-negative by the reference's assert */
        int *hz = &horizon[(section * qadv + qs % qadv) * tadv + ts % tadv];
        if (ts < *hz) continue;
        *hz = ext[k].target_end;
        if (ext[k].stored) out[n++] = ext[k];
    }
    free(horizon);
    return n;
}

/* ---- Heuristic_Span_integrate (src/bsdp/heuristic.c:589-678) ------------------------------
 * For every cell (i,j) of the dst region: the src-region cell with the best START-side score
 * among those a span of [min_query,max_query] x [min_target,max_target] symbols can bridge,
 * first in (query, target) scan order on ties (strict '<', :638), or (-1,-1) when the window is
 * empty.  Heuristic_Span_score is the constant 0 (:362-366), so the reference's re-use of the
 * previous cell's answer when the window did not move (:618-621) changes nothing but time; it
 * is restated anyway so that the walk is the reference's. */
void c4o_span_integrate(const int32_t *src_scores, const int32_t *src_region, const int32_t *dst_region,
                        const int32_t *span, int32_t *out) {
    const int sqs = src_region[0], sts = src_region[1], sql = src_region[2], stl = src_region[3];
    const int dqs = dst_region[0], dts = dst_region[1], dql = dst_region[2], dtl = dst_region[3];
    int prev_iq = -1, prev_fq = -1, prev_it = -1, prev_ft = -1;
    int top_q = -1, top_t = -1;
    int32_t top = 0;
    for (int i = 0; i <= dql; i++)
        for (int j = 0; j <= dtl; j++) {
            int iq = dqs + i - span[1], it = dts + j - span[3];
            int fq = dqs + i - span[0], ft = dts + j - span[2];
            if (iq < sqs) iq = sqs;
            if (it < sts) it = sts;
            if (fq > sqs + sql) fq = sqs + sql;
            if (ft > sts + stl) ft = sts + stl;
            if (iq != prev_iq || it != prev_it || fq != prev_fq || ft != prev_ft) {
                top = C4B_IMPOSSIBLY_LOW_SCORE;
                top_q = top_t = -1;
                for (int x = iq; x <= fq; x++)
                    for (int y = it; y <= ft; y++) {
                        const int32_t cand = src_scores[(size_t)(x - sqs) * (stl + 1) + (y - sts)];
                        if (top < cand) { top = cand; top_q = x; top_t = y; }
                    }
            }
            out[2 * ((size_t)i * (dtl + 1) + j)] = top_q;
            out[2 * ((size_t)i * (dtl + 1) + j) + 1] = top_t;
            prev_iq = iq; prev_it = it; prev_fq = fq; prev_ft = ft;
        }
}
