/* ref_driver.c -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * A thin in-process driver around the UNMODIFIED reference C4 library
 * (compiled in place from /root/reference by oracle/Makefile into
 * oracle/_ref/libc4ref.so).  It is our own code; it only calls the reference's
 * public C API:
 *   Model_Type_get_model / Model_Type_create_data   src/model/modeltype.h:55-62
 *   Optimal_create / Optimal_find_score / _find_path src/c4/optimal.h:49-65
 *   SubOpt_create / SubOpt_add_alignment            src/c4/subopt.h:41-44
 *   Alignment_display_vulgar / _cigar               src/c4/alignment.h:66-72
 * Used to (1) dump closed C4_Models (transition order = tie-break contract),
 * (2) generate golden score/path/vulgar vectors committed under tests/golden/,
 * (3) time the reference's CPU path (bench.py --impl reference / cpu_baseline).
 *
 * The reference owns main() (src/general/argument.c:319); the Makefile renames
 * it to c4ref_unused_main so that argument parsing (which initialises all the
 * function-static ArgumentSets, e.g. affine.c:19-51) runs exactly as in the
 * exonerate binary.  c4ref_session() enters it and calls back.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <string.h>
#include <time.h>

#include "modeltype.h"
#include "optimal.h"
#include "alignment.h"
#include "subopt.h"
#include "viterbi.h"
#include "codegen.h"
#include "heuristic.h"
#include "sdp.h"
#include "bsdp.h"
#include "match.h"
#include "affine.h"
#include "ner.h"
#include "intron.h"
#include "frameshift.h"
#include "alphabet.h"
#include "hspset.h"
#include "sar.h"
#include "splice.h"
#include "translate.h"
#include "sequence.h"

int c4ref_unused_main(int argc, char **argv);

typedef int (*c4ref_callback)(void *ctx);
static c4ref_callback session_cb = NULL;
static void *session_ctx = NULL;
static int session_rc = 0;

int Argument_main(Argument *arg) {
    Translate_ArgumentSet_create(arg);
    Viterbi_ArgumentSet_create(arg);
    Codegen_ArgumentSet_create(arg);
    Heuristic_ArgumentSet_create(arg);
    SDP_ArgumentSet_create(arg);
    BSDP_ArgumentSet_create(arg);
    Sequence_ArgumentSet_create(arg);
    Match_ArgumentSet_create(arg);
    Affine_ArgumentSet_create(arg);
    NER_ArgumentSet_create(arg);
    Intron_ArgumentSet_create(arg);
    Frameshift_ArgumentSet_create(arg);
    Alphabet_ArgumentSet_create(arg);
    HSPset_ArgumentSet_create(arg);
    Alignment_ArgumentSet_create(arg);
    SAR_ArgumentSet_create(arg);
    Splice_ArgumentSet_create(arg);
    Argument_process(arg, "c4ref", "in-process reference driver", NULL);
    session_rc = session_cb ? session_cb(session_ctx) : 0;
    return session_rc;
}

/* Runs cb(ctx) inside a fully initialised reference process context.
 * argv[0] is the program name; the rest are exonerate options such as
 * "--gapopen", "-12", "--dpmemory", "32", "--compiled", "no". */
int c4ref_session(int argc, char **argv, c4ref_callback cb, void *ctx) {
    session_cb = cb;
    session_ctx = ctx;
    return c4ref_unused_main(argc, argv);
}

/* ---------------------------------------------------------------------- */

typedef struct {
    Model_Type type;
    Alphabet_Type qtype, ttype;
    C4_Model *model;
    Optimal *optimal;
    Alphabet *qalpha, *talpha;
} c4ref_model;

void *c4ref_model_open(const char *model_name, int query_is_protein,
                       int target_is_protein, int use_compiled) {
    c4ref_model *m = g_new0(c4ref_model, 1);
    m->type = Model_Type_from_string((gchar *)model_name);
    m->qtype = query_is_protein ? Alphabet_Type_PROTEIN : Alphabet_Type_DNA;
    m->ttype = target_is_protein ? Alphabet_Type_PROTEIN : Alphabet_Type_DNA;
    m->model = Model_Type_get_model(m->type, m->qtype, m->ttype);
    /* same flags as GAM_create, src/hub/gam.c:420-424 */
    m->optimal = Optimal_create(m->model, NULL,
                                Optimal_Type_SCORE | Optimal_Type_PATH |
                                    Optimal_Type_REDUCED_SPACE,
                                use_compiled ? TRUE : FALSE);
    m->qalpha = Alphabet_create(m->qtype, FALSE);
    m->talpha = Alphabet_create(m->ttype, FALSE);
    return m;
}

void c4ref_model_close(void *h) {
    c4ref_model *m = (c4ref_model *)h;
    Optimal_destroy(m->optimal);
    C4_Model_destroy(m->model);
    Alphabet_destroy(m->qalpha);
    Alphabet_destroy(m->talpha);
    g_free(m);
}

static void dump_str(FILE *fp, const char *key, const char *s) {
    fprintf(fp, " %s=\"", key);
    if (s)
        for (; *s; s++) fputc((*s == '"' || *s == '\n') ? ' ' : *s, fp);
    fputc('"', fp);
}

/* Text dump of the CLOSED model: one record per line. Caller frees with
 * c4ref_free(). */
char *c4ref_model_dump(void *h) {
    c4ref_model *m = (c4ref_model *)h;
    C4_Model *model = m->model;
    char *buf = NULL;
    size_t len = 0;
    FILE *fp = open_memstream(&buf, &len);
    guint i, j;
    fprintf(fp, "model");
    dump_str(fp, "name", model->name);
    fprintf(fp,
            " states=%u transitions=%u calcs=%u shadows=%u portals=%u spans=%u"
            " max_query_advance=%d max_target_advance=%d shadow_designations=%d"
            " start_state=%d start_scope=%d end_state=%d end_scope=%d"
            " start_cell_func=%d end_cell_func=%d\n",
            model->state_list->len, model->transition_list->len,
            model->calc_list->len, model->shadow_list->len,
            model->portal_list->len, model->span_list->len,
            model->max_query_advance, model->max_target_advance,
            model->total_shadow_designations, model->start_state->state->id,
            (int)model->start_state->scope, model->end_state->state->id,
            (int)model->end_state->scope,
            model->start_state->cell_start_func ? 1 : 0,
            model->end_state->cell_end_func ? 1 : 0);
    for (i = 0; i < model->state_list->len; i++) {
        C4_State *s = model->state_list->pdata[i];
        fprintf(fp, "state id=%d", s->id);
        dump_str(fp, "name", s->name);
        fprintf(fp, " src_shadows=");
        for (j = 0; j < s->src_shadow_list->len; j++) {
            C4_Shadow *sh = s->src_shadow_list->pdata[j];
            fprintf(fp, "%s%d", j ? "," : "", sh->id);
        }
        fprintf(fp, "\n");
    }
    for (i = 0; i < model->calc_list->len; i++) {
        C4_Calc *c = model->calc_list->pdata[i];
        fprintf(fp, "calc id=%d", c->id);
        dump_str(fp, "name", c->name);
        fprintf(fp, " max_score=%d protect=%d has_func=%d has_init=%d has_exit=%d",
                c->max_score, (int)c->protect, c->calc_func ? 1 : 0,
                c->init_func ? 1 : 0, c->exit_func ? 1 : 0);
        dump_str(fp, "macro", c->calc_macro);
        fprintf(fp, "\n");
    }
    for (i = 0; i < model->transition_list->len; i++) {
        C4_Transition *t = model->transition_list->pdata[i];
        fprintf(fp, "transition id=%d", t->id);
        dump_str(fp, "name", t->name);
        fprintf(fp, " input=%d output=%d advance_query=%d advance_target=%d calc=%d label=%d",
                t->input->id, t->output->id, t->advance_query, t->advance_target,
                t->calc ? t->calc->id : -1, (int)t->label);
        fprintf(fp, " dst_shadows=");
        for (j = 0; j < t->dst_shadow_list->len; j++) {
            C4_Shadow *sh = t->dst_shadow_list->pdata[j];
            fprintf(fp, "%s%d", j ? "," : "", sh->id);
        }
        fprintf(fp, "\n");
    }
    for (i = 0; i < model->shadow_list->len; i++) {
        C4_Shadow *sh = model->shadow_list->pdata[i];
        fprintf(fp, "shadow id=%d", sh->id);
        dump_str(fp, "name", sh->name);
        fprintf(fp, " designation=%d src_states=", sh->designation);
        for (j = 0; j < sh->src_state_list->len; j++) {
            C4_State *s = sh->src_state_list->pdata[j];
            fprintf(fp, "%s%d", j ? "," : "", s->id);
        }
        fprintf(fp, " dst_transitions=");
        for (j = 0; j < sh->dst_transition_list->len; j++) {
            C4_Transition *t = sh->dst_transition_list->pdata[j];
            fprintf(fp, "%s%d", j ? "," : "", t->id);
        }
        dump_str(fp, "start_macro", sh->start_macro);
        dump_str(fp, "end_macro", sh->end_macro);
        fprintf(fp, "\n");
    }
    for (i = 0; i < model->portal_list->len; i++) {
        C4_Portal *p = model->portal_list->pdata[i];
        fprintf(fp, "portal id=%d", p->id);
        dump_str(fp, "name", p->name);
        fprintf(fp, " advance_query=%d advance_target=%d calc=%d\n",
                p->advance_query, p->advance_target, p->calc ? p->calc->id : -1);
    }
    for (i = 0; i < model->span_list->len; i++) {
        C4_Span *sp = model->span_list->pdata[i];
        fprintf(fp, "span id=%d", sp->id);
        dump_str(fp, "name", sp->name);
        fprintf(fp, " state=%d min_query=%d max_query=%d min_target=%d max_target=%d\n",
                sp->span_state->id, sp->min_query, sp->max_query, sp->min_target,
                sp->max_target);
    }
    fclose(fp);
    return buf;
}

void c4ref_free(void *p) { free(p); }

/* ---------------------------------------------------------------------- */

typedef struct {
    c4ref_model *m;
    Sequence *query, *target;
    gpointer user_data;
    SubOpt *subopt;
    Region *region;
} c4ref_pair;

void *c4ref_pair_open(void *h, const char *qid, const char *qseq,
                      const char *tid, const char *tseq) {
    c4ref_model *m = (c4ref_model *)h;
    c4ref_pair *p = g_new0(c4ref_pair, 1);
    p->m = m;
    p->query = Sequence_create((gchar *)qid, NULL, (gchar *)qseq, 0,
                               (m->qtype == Alphabet_Type_DNA)
                                   ? Sequence_Strand_FORWARD
                                   : Sequence_Strand_UNKNOWN,
                               m->qalpha);
    p->target = Sequence_create((gchar *)tid, NULL, (gchar *)tseq, 0,
                                (m->ttype == Alphabet_Type_DNA)
                                    ? Sequence_Strand_FORWARD
                                    : Sequence_Strand_UNKNOWN,
                                m->talpha);
    p->user_data = Model_Type_create_data(m->type, p->query, p->target);
    p->subopt = SubOpt_create(p->query->len, p->target->len);
    p->region = Region_create(0, 0, p->query->len, p->target->len);
    return p;
}

void c4ref_pair_close(void *ph) {
    c4ref_pair *p = (c4ref_pair *)ph;
    Region_destroy(p->region);
    SubOpt_destroy(p->subopt);
    Model_Type_destroy_data(p->m->type, p->user_data);
    Sequence_destroy(p->query);
    Sequence_destroy(p->target);
    g_free(p);
}

int c4ref_pair_score(void *ph, int use_subopt) {
    c4ref_pair *p = (c4ref_pair *)ph;
    return Optimal_find_score(p->m->optimal, p->region, p->user_data,
                              use_subopt ? p->subopt : NULL);
}

/* One Optimal_find_path() call, as OPair_next_path (src/c4/opair.c:42-56).
 * If add_to_subopt, the alignment is then recorded in the pair's SubOpt
 * (GAM_Result_add_alignment -> SubOpt_add_alignment, src/hub/gam.c) so the
 * next call returns the next sub-optimal alignment.
 * Outputs: region4 = {query_start, target_start, query_length, target_length};
 * ops = (transition_id, length) pairs; *vulgar / *cigar malloc'd lines.
 * Returns 1 if an alignment was produced, 0 if below threshold. */
int c4ref_pair_path(void *ph, int threshold, int use_subopt, int add_to_subopt,
                    int *score, int *region4, int *ops, int max_ops, int *n_ops,
                    char **vulgar, char **cigar) {
    c4ref_pair *p = (c4ref_pair *)ph;
    Alignment *a = Optimal_find_path(p->m->optimal, p->region, p->user_data,
                                     threshold, use_subopt ? p->subopt : NULL);
    guint i;
    if (!a) return 0;
    *score = a->score;
    region4[0] = a->region->query_start;
    region4[1] = a->region->target_start;
    region4[2] = a->region->query_length;
    region4[3] = a->region->target_length;
    *n_ops = (int)a->operation_list->len;
    for (i = 0; i < a->operation_list->len && (int)i < max_ops; i++) {
        AlignmentOperation *ao = a->operation_list->pdata[i];
        ops[2 * i] = ao->transition->id;
        ops[2 * i + 1] = ao->length;
    }
    if (vulgar) {
        size_t len = 0;
        FILE *fp = open_memstream(vulgar, &len);
        Alignment_display_vulgar(a, p->query, p->target, fp);
        fclose(fp);
    }
    if (cigar) {
        size_t len = 0;
        FILE *fp = open_memstream(cigar, &len);
        Alignment_display_cigar(a, p->query, p->target, fp);
        fclose(fp);
    }
    if (add_to_subopt) SubOpt_add_alignment(p->subopt, a);
    Alignment_destroy(a);
    return 1;
}

/* Wall-clock helper for the CPU baseline: n_rep full find_path calls. */
double c4ref_pair_time_path(void *ph, int n_rep, int *score) {
    c4ref_pair *p = (c4ref_pair *)ph;
    struct timespec t0, t1;
    int r;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (r = 0; r < n_rep; r++) {
        Alignment *a = Optimal_find_path(p->m->optimal, p->region, p->user_data,
                                         C4_IMPOSSIBLY_LOW_SCORE, NULL);
        *score = a->score;
        Alignment_destroy(a);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* Scoring primitives, for pinning the oracle's tables. */
int c4ref_submat_lookup(int protein, int a, int b) {
    Match_ArgumentSet *mas = Match_ArgumentSet_create(NULL);
    Submat *s = protein ? mas->protein_submat : mas->dna_submat;
    return Submat_lookup(s, a, b);
}

/* ---- export of the reference's scoring tables (for golden fixtures) ----- */
void c4ref_get_submat(int protein, int *matrix576, unsigned char *index256) {
    Match_ArgumentSet *mas = Match_ArgumentSet_create(NULL);
    Submat *s = protein ? mas->protein_submat : mas->dna_submat;
    int a, b;
    for (a = 0; a < SUBMAT_ALPHABETSIZE; a++)
        for (b = 0; b < SUBMAT_ALPHABETSIZE; b++)
            matrix576[a * SUBMAT_ALPHABETSIZE + b] = s->matrix[a][b];
    for (a = 0; a < 256; a++) index256[a] = s->index[a];
}

void c4ref_get_translate(unsigned char *nt2d256, unsigned char *codon_aa4096) {
    Match_ArgumentSet *mas = Match_ArgumentSet_create(NULL);
    Translate *t = mas->translate;
    int i;
    for (i = 0; i < 256; i++) nt2d256[i] = t->nt2d[i];
    for (i = 0; i < 4096; i++) codon_aa4096[i] = t->aa[t->trans[i]];
}

void c4ref_get_intron_params(int *min_intron, int *max_intron, int *open_penalty) {
    Intron_ArgumentSet *ias = Intron_ArgumentSet_create(NULL);
    *min_intron = ias->min_intron;
    *max_intron = ias->max_intron;
    *open_penalty = ias->intron_open_penalty;
}

void c4ref_get_gap_params(int *gap_open, int *gap_extend, int *codon_gap_open,
                          int *codon_gap_extend, int *frameshift) {
    Affine_ArgumentSet *aas = Affine_ArgumentSet_create(NULL);
    Frameshift_ArgumentSet *fas = Frameshift_ArgumentSet_create(NULL);
    *gap_open = aas->gap_open;
    *gap_extend = aas->gap_extend;
    *codon_gap_open = aas->codon_gap_open;
    *codon_gap_extend = aas->codon_gap_extend;
    *frameshift = fas->frameshift_penalty;
}

/* SplicePredictor_predict_array_int over a whole sequence
 * (src/sequence/splice.c:383-397); type = SpliceType enum value. */
void c4ref_splice_array(int type, const char *seq, int len, int *out) {
    Intron_ArgumentSet *ias = Intron_ArgumentSet_create(NULL);
    SplicePredictor *sp = NULL;
    switch (type) {
    case SpliceType_ss5_forward: sp = ias->sps->ss5_forward; break;
    case SpliceType_ss3_forward: sp = ias->sps->ss3_forward; break;
    case SpliceType_ss5_reverse: sp = ias->sps->ss5_reverse; break;
    case SpliceType_ss3_reverse: sp = ias->sps->ss3_reverse; break;
    }
    SplicePredictor_predict_array_int(sp, (gchar *)seq, (guint)len, 0, (guint)len, out);
}

/* ---- HSP seeding / extension (src/comparison/hspset.c:725-815, 933-997) ----------
 * HSPset_seed_hsp for every (query_start, target_start) in list order, then
 * HSPset_finalise; the pattern of src/comparison/hspset.test.c.  softmask_* selects
 * a soft-masked Alphabet for that sequence (lower case = masked).
 * out: 5 ints per stored HSP {query_start, target_start, length, score, cobs};
 * params: {seedlen, dropoff, threshold, query_advance, target_advance, filter_threshold,
 * seed_repeat}.  Returns the number of HSPs (or -1 when out is too small). */
int c4ref_hspset(int match_type, const char *qseq, const char *tseq, int softmask_q, int softmask_t,
                 const unsigned *seeds, int n_seeds, int *out, int max_out, int *params) {
    Match *match = Match_find((Match_Type)match_type);
    Alphabet *qa = Alphabet_create(match->query->alphabet->type, softmask_q ? TRUE : FALSE);
    Alphabet *ta = Alphabet_create(match->target->alphabet->type, softmask_t ? TRUE : FALSE);
    Sequence *query = Sequence_create("qy", NULL, (gchar *)qseq, 0, Sequence_Strand_UNKNOWN, qa);
    Sequence *target = Sequence_create("tg", NULL, (gchar *)tseq, 0, Sequence_Strand_UNKNOWN, ta);
    HSP_Param *hsp_param = HSP_Param_create(match, TRUE);
    HSPset *hsp_set = HSPset_create(query, target, hsp_param);
    int i, n;
    params[0] = hsp_param->seedlen; params[1] = hsp_param->dropoff; params[2] = hsp_param->threshold;
    params[3] = match->query->advance; params[4] = match->target->advance;
    params[5] = hsp_param->has->filter_threshold; params[6] = hsp_param->seed_repeat;
    for (i = 0; i < n_seeds; i++)
        HSPset_seed_hsp(hsp_set, seeds[2 * i], seeds[2 * i + 1]);
    HSPset_finalise(hsp_set);
    n = hsp_set->hsp_list->len;
    if (n > max_out) n = -1;
    for (i = 0; i < n; i++) {
        HSP *hsp = hsp_set->hsp_list->pdata[i];
        out[5 * i + 0] = hsp->query_start; out[5 * i + 1] = hsp->target_start;
        out[5 * i + 2] = hsp->length; out[5 * i + 3] = hsp->score; out[5 * i + 4] = hsp->cobs;
    }
    HSPset_destroy(hsp_set);
    HSP_Param_destroy(hsp_param);
    Sequence_destroy(query);
    Sequence_destroy(target);
    Alphabet_destroy(qa);
    Alphabet_destroy(ta);
    return n;
}

/* ---- Heuristic_Span_integrate (src/bsdp/heuristic.c:589-678) on caller-supplied matrices ----
 * The reference function reads only span->{min,max}_{query,target}, src_integration_matrix
 * [x][y][0] and the two regions, and writes dst_integration_matrix[i][j].{query,target}_pos:
 * a Heuristic_Span holding just those is enough to run the UNMODIFIED function.
 * src_scores: (src_ql+1) x (src_tl+1) ints; regions {q_start, t_start, q_len, t_len};
 * span {min_query, max_query, min_target, max_target};
 * out: (dst_ql+1) x (dst_tl+1) x {query_pos, target_pos}. */
void c4ref_span_integrate(const int *src_scores, const int *src_region, const int *dst_region,
                          const int *span, int *out) {
    Heuristic_Span hs;
    C4_Span sp;
    Region *src = Region_create(src_region[0], src_region[1], src_region[2], src_region[3]);
    Region *dst = Region_create(dst_region[0], dst_region[1], dst_region[2], dst_region[3]);
    int i, j, sql = src_region[2], stl = src_region[3], dql = dst_region[2], dtl = dst_region[3];
    memset(&hs, 0, sizeof(hs));
    memset(&sp, 0, sizeof(sp));
    sp.min_query = span[0]; sp.max_query = span[1]; sp.min_target = span[2]; sp.max_target = span[3];
    hs.span = &sp;
    hs.src_integration_matrix = g_new(C4_Score **, sql + 1);
    for (i = 0; i <= sql; i++) {
        hs.src_integration_matrix[i] = g_new(C4_Score *, stl + 1);
        for (j = 0; j <= stl; j++) {
            hs.src_integration_matrix[i][j] = g_new(C4_Score, 1);
            hs.src_integration_matrix[i][j][0] = src_scores[i * (stl + 1) + j];
        }
    }
    hs.dst_integration_matrix = g_new(Heuristic_Span_Cell *, dql + 1);
    for (i = 0; i <= dql; i++) hs.dst_integration_matrix[i] = g_new0(Heuristic_Span_Cell, dtl + 1);
    Heuristic_Span_integrate(&hs, src, dst);
    for (i = 0; i <= dql; i++)
        for (j = 0; j <= dtl; j++) {
            out[2 * (i * (dtl + 1) + j)] = hs.dst_integration_matrix[i][j].query_pos;
            out[2 * (i * (dtl + 1) + j) + 1] = hs.dst_integration_matrix[i][j].target_pos;
        }
    for (i = 0; i <= sql; i++) {
        for (j = 0; j <= stl; j++) g_free(hs.src_integration_matrix[i][j]);
        g_free(hs.src_integration_matrix[i]);
    }
    g_free(hs.src_integration_matrix);
    for (i = 0; i <= dql; i++) g_free(hs.dst_integration_matrix[i]);
    g_free(hs.dst_integration_matrix);
    Region_destroy(src);
    Region_destroy(dst);
}
