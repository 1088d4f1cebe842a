/* viterbi_b200.c -- the reference-side binding of libc4b200.so.
 *
 * A drop-in replacement for the reference's src/c4/viterbi.c: same public
 * functions (src/c4/viterbi.h:122-159), but the lattice fill and the traceback
 * run on the GPU through the C ABI of include/c4b200.h.  It is linked INTO THE
 * UNMODIFIED REFERENCE in place of viterbi.o (the substitution point the
 * reference's own build provides, src/program/Makefile.am:18,76-86); nothing in
 * src/model, src/hub, optimal.c or alignment.c changes.  See INTEGRATION.md.
 *
 * Our own code; it includes the reference's headers because it implements the
 * reference's interface.  Built only where /root/reference exists
 * (integration/Makefile) into integration/_build/exonerate_b200.
 */
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <time.h>

#include "viterbi.h"
#include "ungapped.h"
#include "affine.h"
#include "intron.h"
#include "frameshift.h"
#include "match.h"
#include "splice.h"
#include "c4b200.h"
#include "heuristic.h"
#include "b200_binding.h"

/* ---- options (viterbi.c:27-38): -D/--dpmemory is accepted and ignored -------- */
Viterbi_ArgumentSet *Viterbi_ArgumentSet_create(Argument *arg){
    register ArgumentSet *as;
    static Viterbi_ArgumentSet vas = {32};
    if(arg){
        as = ArgumentSet_create("Viterbi algorithm options");
        ArgumentSet_add_option(as, 'D', "dpmemory", "Mb",
           "Maximum memory to use for DP tracebacks (Mb) [unused: GPU engine]",
           "32", Argument_parse_int, &vas.traceback_memory_limit);
        Argument_absorb_ArgumentSet(arg, as);
        }
    return &vas;
    }

/* ---- engine + per-Viterbi tables --------------------------------------------- */
static c4b_engine *engine = NULL;

#define get_engine exonerate_b200_engine /* shared with hspset_b200.c, gam_b200.c */

B200_Replay *b200_replay = NULL;
glong b200_stat_prefetch_hits = 0, b200_stat_prefetch_misses = 0;
glong b200_stat_score_hits = 0, b200_stat_score_prefetched = 0, b200_stat_score_batches = 0;
static void score_cache_clear(void);

c4b_engine *exonerate_b200_engine(void){
    if(!engine){
        register const gchar *dev = g_getenv("EXONERATE_B200_DEVICE");
        if(c4b_engine_create(dev?atoi(dev):0, &engine))
            g_error("libc4b200: %s", c4b_last_error());
        }
    return engine;
    }

c4b_group *exonerate_b200_group(void){
    static c4b_group *group = NULL;
    static gboolean tried = FALSE;
    register const gchar *spec = g_getenv("EXONERATE_B200_DEVICES");
    int devices[64], n = 0;
    if(tried || !spec || !spec[0])
        return group;
    tried = TRUE;
    if(strcmp(spec, "all")){
        register gchar **part = g_strsplit(spec, ",", 64);
        for(n = 0; part[n] && (n < 64); n++)
            devices[n] = atoi(part[n]);
        g_strfreev(part);
        }
    if(c4b_group_create(n, devices, &group))
        g_error("libc4b200: %s", c4b_last_error());
    return group;
    }

/* EXONERATE_B200_STATS=1: one line on stderr at exit */
static glong stat_calls = 0, stat_cache_miss = 0;
static gdouble stat_total = 0, stat_prepare = 0, stat_engine = 0, stat_prefetch = 0, stat_spans = 0;
static gdouble now_seconds(void){
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9*ts.tv_nsec;
    }
/* EXONERATE_B200_TRACE=1: one line per Viterbi_calculate on stderr (debugging aid: diff two runs) */
static void trace_call(Viterbi *viterbi, Region *region, const gchar *how, C4_Score score){
    static gint on = -1;
    if(on < 0)
        on = g_getenv("EXONERATE_B200_TRACE")?1:0;
    if(on)
        fprintf(stderr, "b200-trace %s mode %d [%s] region %d %d %d %d -> %d\n", how,
                (gint)viterbi->mode, viterbi->name, region->query_start, region->target_start,
                region->query_length, region->target_length, score);
    return;
    }

static void print_viterbi_stats(void){
    fprintf(stderr, "exonerate_b200: Viterbi_calculate calls %ld (%.3f s: prepare %.3f s, "
                    "engine %.3f s), targets flattened %ld, answered from the batch "
                    "prefetch %ld (prefetched but not usable %ld); BSDP region fills prefetched %ld "
                    "in %ld batch(es), answered from them %ld (terminal / join batches %.3f s, span "
                    "batches %.3f s)\n",
            stat_calls, stat_total, stat_prepare, stat_engine, stat_cache_miss,
            b200_stat_prefetch_hits, b200_stat_prefetch_misses,
            b200_stat_score_prefetched, b200_stat_score_batches, b200_stat_score_hits,
            stat_prefetch, stat_spans);
    }

/* Flattened sequences (they are virtual in the reference: revcomp / subseq / translate
 * views, sequence.c:257-507) and the four splice-site arrays intron_init_func prepares
 * (intron.c:259-293) of the CURRENT comparison.  BSDP asks for thousands of region fills
 * on the same (query, target): flatten once per pair, not once per fill.  The Sequences
 * are shared (pinned) while cached, so the pointers cannot be recycled under the key.
 * TARGETS are kept across comparisons, by content: the usual run compares many queries (and
 * their reverse complements) with the same genomic targets, which the reference re-reads --
 * as new Sequence objects -- for every query (src/database/fastapipe.c:106-137).  Flattening a
 * 6 Mbp target and predicting its four splice arrays per comparison cost 134 s of a 242 s run
 * (profiles/r02_bsdp.md); kept, they are computed once, and the engine's device copies
 * (C4B_PAIR_BUFFERS_STABLE, keyed by the host address) stay valid with them. */
typedef struct {
    guint64 hash;
    gint len;
    gchar *tseq;
    gint32 *splice;
    glong last_use;
} B200_Target;
#define TARGET_SLOTS 8
static B200_Target target_cache[TARGET_SLOTS];
static glong target_clock = 0;
static gsize target_cache_bytes = 0;

static struct {
    Sequence *query, *target;
    gchar *qseq, *tseq;
    gint32 *splice;
    B200_Target *entry;
} pair_cache = {NULL, NULL, NULL, NULL, NULL, NULL};

static guint64 content_hash(const gchar *s, gint len){
    register guint64 h = 0x9E3779B97F4A7C15ull ^ (guint64)len;
    register gint i;
    guint64 w;
    for(i = 0; i+8 <= len; i += 8){
        memcpy(&w, s+i, 8);
        h = (h ^ w) * 0xFF51AFD7ED558CCDull;
        h ^= h >> 32;
        }
    for(; i < len; i++)
        h = (h ^ (guchar)s[i]) * 0x100000001B3ull;
    return h ^ (h >> 29);
    }

static void target_entry_drop(B200_Target *e){
    if(!e->tseq)
        return;
    c4b_engine_forget_buffer(get_engine(), e->tseq); /* device copies are keyed by these addresses */
    if(e->splice){
        register gint k;
        for(k = 0; k < 4; k++)
            c4b_engine_forget_buffer(get_engine(), e->splice + (gsize)k*e->len);
        target_cache_bytes -= 16*((gsize)e->len+1);
        }
    target_cache_bytes -= e->len;
    g_free(e->tseq);
    g_free(e->splice);
    memset(e, 0, sizeof(B200_Target));
    return;
    }

static B200_Target *target_fetch(Sequence *target){
    register gint i, tlen = target->len;
    register gchar *flat = g_new(gchar, tlen+4);
    register guint64 h;
    register B200_Target *e, *victim = NULL;
    Sequence_strncpy(target, 0, tlen, flat);
    memset(flat+tlen, 0, 4);
    h = content_hash(flat, tlen);
    for(i = 0; i < TARGET_SLOTS; i++){
        e = &target_cache[i];
        if(e->tseq && (e->hash == h) && (e->len == tlen) && !memcmp(e->tseq, flat, tlen)){
            g_free(flat);
            e->last_use = ++target_clock;
            return e;
            }
        }
    /* keep at most TARGET_SLOTS targets / 2 GB of sequence + splice arrays: drop the oldest */
    for(;;){
        register B200_Target *oldest = NULL;
        victim = NULL;
        for(i = 0; i < TARGET_SLOTS; i++){
            e = &target_cache[i];
            if(!e->tseq)
                victim = e;
            else if((e != pair_cache.entry) && (!oldest || (e->last_use < oldest->last_use)))
                oldest = e;
            }
        if(victim && ((target_cache_bytes + 17*(gsize)tlen <= ((gsize)2 << 30)) || !oldest))
            break;
        if(!oldest)
            g_error("libc4b200: target cache exhausted");
        target_entry_drop(oldest);
        }
    victim->hash = h;
    victim->len = tlen;
    victim->tseq = flat;
    victim->splice = NULL;
    victim->last_use = ++target_clock;
    target_cache_bytes += tlen;
    stat_cache_miss++;
    return victim;
    }

gint32 *b200_splice_arrays(gchar *tseq, gint tlen){
    register Intron_ArgumentSet *ias = Intron_ArgumentSet_create(NULL);
    register gint32 *splice = g_new(gint32, 4*((gsize)tlen+1));
    SplicePredictor_predict_array_int(ias->sps->ss5_forward, tseq, tlen, 0, tlen, splice);
    SplicePredictor_predict_array_int(ias->sps->ss3_forward, tseq, tlen, 0, tlen, splice+tlen);
    SplicePredictor_predict_array_int(ias->sps->ss5_reverse, tseq, tlen, 0, tlen,
                                      splice+2*(gsize)tlen);
    SplicePredictor_predict_array_int(ias->sps->ss3_reverse, tseq, tlen, 0, tlen,
                                      splice+3*(gsize)tlen);
    return splice;
    }

static void pair_cache_fetch(Ungapped_Data *ud, gboolean need_splice){
    register gint tlen = ud->target->len;
    if((pair_cache.query != ud->query) || (pair_cache.target != ud->target)){
        score_cache_clear(); /* prefetched BSDP fills belong to the previous comparison */
        if(pair_cache.query != ud->query){
            if(pair_cache.query){
                c4b_engine_forget_buffer(get_engine(), pair_cache.qseq);
                Sequence_destroy(pair_cache.query);
                g_free(pair_cache.qseq);
                }
            pair_cache.query = Sequence_share(ud->query);
            pair_cache.qseq = g_new(gchar, ud->query->len+4);
            Sequence_strncpy(ud->query, 0, ud->query->len, pair_cache.qseq);
            }
        if(pair_cache.target != ud->target){
            if(pair_cache.target)
                Sequence_destroy(pair_cache.target);
            pair_cache.target = Sequence_share(ud->target);
            pair_cache.entry = NULL;
            pair_cache.entry = target_fetch(ud->target);
            pair_cache.tseq = pair_cache.entry->tseq;
            pair_cache.splice = pair_cache.entry->splice;
            }
        }
    if(need_splice && !pair_cache.splice){
        pair_cache.entry->splice = b200_splice_arrays(pair_cache.tseq, tlen);
        pair_cache.splice = pair_cache.entry->splice;
        target_cache_bytes += 16*((gsize)tlen+1);
        }
    return;
    }

typedef struct B200_Tables {
    Viterbi *viterbi;
    c4b_model model;
    struct B200_Tables *next;
} B200_Tables;
static B200_Tables *tables_list = NULL;

typedef struct { /* what one FIND_PATH call leaves behind for create_Alignment */
    c4b_result result;
    gint32 *ops;
} B200_Path;

/* Map one C4_Calc of the reference to a device calc kind by its calc_macro
 * (the text the reference's own code generator pastes; every calc of every
 * compiled model has one).  Unknown => g_error: there is no host fallback. */
static void classify_calc(C4_Model *model, C4_Calc *calc, c4b_calc *out){
    register gchar *m = calc->calc_macro;
    register Affine_ArgumentSet *aas = Affine_ArgumentSet_create(NULL);
    register Frameshift_ArgumentSet *fas = Frameshift_ArgumentSet_create(NULL);
    register Intron_ArgumentSet *ias = Intron_ArgumentSet_create(NULL);
    memset(out, 0, sizeof(c4b_calc));
    out->protect = calc->protect;
    if(!calc->calc_func){ /* C4_Calc_score (c4.c:1700-1711): no callback => max_score;
                           * that is every calc of a BSDP bound model (heuristic.c:172-180) */
        out->kind = C4B_CALC_CONST;
        out->param[0] = calc->max_score;
        return;
        }
    if(!m)
        g_error("libc4b200: calc [%s] of model [%s] has no macro to classify",
                calc->name, model->name);
    if(strstr(m, "split_score_func")){
        out->kind = strstr(m, "curr_intron_start >= 1")
                  ? C4B_CALC_PHASE1_POST : C4B_CALC_PHASE2_POST;
    } else if(strstr(m, "SplicePrediction_get")){
        if(strstr(m, "sps->ss5_forward")) out->param[1] = C4B_SPLICE_5_FORWARD;
        else if(strstr(m, "sps->ss3_forward")) out->param[1] = C4B_SPLICE_3_FORWARD;
        else if(strstr(m, "sps->ss5_reverse")) out->param[1] = C4B_SPLICE_5_REVERSE;
        else if(strstr(m, "sps->ss3_reverse")) out->param[1] = C4B_SPLICE_3_REVERSE;
        else g_error("libc4b200: unknown splice site in calc [%s]", calc->name);
        if(strstr(m, "query_data"))
            g_error("libc4b200: query introns have no device form [%s]", calc->name);
        if(strstr(m, "intron_open_penalty")){
            out->kind = C4B_CALC_SPLICE_PRE;
            out->param[0] = ias->intron_open_penalty;
        } else {
            out->kind = C4B_CALC_SPLICE_POST;
            }
    } else if(strstr(m, "Submat_lookup")){
        register gboolean tq = strstr(m, "Sequence_get_symbol(ud->query, %QP+1)")?TRUE:FALSE,
                          tt = strstr(m, "Sequence_get_symbol(ud->target, %TP+1)")?TRUE:FALSE;
        if(strstr(m, "dna_submat")) out->kind = C4B_CALC_MATCH_DNA;
        else if(tq && tt) out->kind = C4B_CALC_MATCH_3_3;
        else if(tt) out->kind = C4B_CALC_MATCH_1_3;
        else if(tq) out->kind = C4B_CALC_MATCH_3_1;
        else out->kind = C4B_CALC_MATCH_PROTEIN;
    } else if(strstr(m, "aas->codon_gap_open)")){
        out->kind = C4B_CALC_CONST; out->param[0] = aas->codon_gap_open;
    } else if(strstr(m, "aas->codon_gap_extend)")){
        out->kind = C4B_CALC_CONST; out->param[0] = aas->codon_gap_extend;
    } else if(strstr(m, "aas->gap_open)")){
        out->kind = C4B_CALC_CONST; out->param[0] = aas->gap_open;
    } else if(strstr(m, "aas->gap_extend)")){
        out->kind = C4B_CALC_CONST; out->param[0] = aas->gap_extend;
    } else if(strstr(m, "frameshift_penalty")){
        out->kind = C4B_CALC_CONST; out->param[0] = fas->frameshift_penalty;
    } else {
        g_error("libc4b200: calc [%s] of model [%s] has no device form",
                calc->name, model->name);
        }
    return;
    }

/* closed C4_Model -> c4b_model */
/* Heuristic_Bound_create (src/bsdp/heuristic.c:150-207) scores a model whose only cell
 * callback is Heuristic_Bound_report_end_func, recognised by its macro (:147-148). */
static gboolean model_is_bound(C4_Model *model){
    return model->end_state->cell_end_func && model->end_state->cell_end_macro
        && !strcmp(model->end_state->cell_end_macro, "matrix[%QP][%TP] = %C[0]");
    }

static void flatten_model(C4_Model *model, c4b_model *out){
    register gint i, j;
    register C4_Transition *t;
    register C4_Shadow *shadow;
    register C4_State *state;
    g_assert(!model->is_open);
    memset(out, 0, sizeof(c4b_model));
    if((model->state_list->len > C4B_MAX_STATES)
    || (model->transition_list->len > C4B_MAX_TRANSITIONS)
    || (model->calc_list->len > C4B_MAX_CALCS)
    || (model->total_shadow_designations > C4B_MAX_SHADOW_SLOTS))
        g_error("libc4b200: model [%s] exceeds the engine's table sizes",
                model->name);
    /* cell_start_func / cell_end_func (BSDP derived models) are evaluated by
     * Viterbi_calculate around the device fill: see c4b_viterbi_calculate_cells */
    out->n_states = model->state_list->len;
    out->n_transitions = model->transition_list->len;
    out->n_calcs = model->calc_list->len;
    out->n_shadow_slots = model->total_shadow_designations;
    out->start_state = model->start_state->state->id;
    out->end_state = model->end_state->state->id;
    out->start_scope = model->start_state->scope;
    out->end_scope = model->end_state->scope;
    out->max_query_advance = model->max_query_advance;
    out->max_target_advance = model->max_target_advance;
    for(i = 0; i < model->calc_list->len; i++)
        classify_calc(model, model->calc_list->pdata[i], &out->calcs[i]);
    for(i = 0; i < model->transition_list->len; i++){
        t = model->transition_list->pdata[i]; /* closed order = tie-break contract */
        out->transitions[i].input = t->input->id;
        out->transitions[i].output = t->output->id;
        out->transitions[i].advance_query = t->advance_query;
        out->transitions[i].advance_target = t->advance_target;
        out->transitions[i].calc = t->calc?t->calc->id:-1;
        out->transitions[i].label = t->label;
        if(t->calc && (out->calcs[t->calc->id].kind >= C4B_CALC_SPLICE_POST)){
            g_assert(t->dst_shadow_list->len == 1);
            shadow = t->dst_shadow_list->pdata[0];
            out->calcs[t->calc->id].param[2] = shadow->designation;
            }
        }
    for(i = 0; i < model->shadow_list->len; i++){
        shadow = model->shadow_list->pdata[i];
        for(j = 0; j < shadow->src_state_list->len; j++){
            state = shadow->src_state_list->pdata[j];
            out->shadow_start[state->id][shadow->designation]
                = strstr(shadow->start_macro, "%TP")?1:2;
            }
        }
    return;
    }

c4b_model *b200_tables_for(Viterbi *viterbi){
    register B200_Tables *bt;
    for(bt = tables_list; bt; bt = bt->next)
        if(bt->viterbi == viterbi)
            return &bt->model;
    bt = g_new0(B200_Tables, 1);
    bt->viterbi = viterbi;
    flatten_model(viterbi->model, &bt->model);
    bt->next = tables_list;
    tables_list = bt;
    return &bt->model;
    }

/* ---- Viterbi objects (viterbi.c:58-104) ----------------------------------------- */
Viterbi *Viterbi_create(C4_Model *model, gchar *name,
                        Viterbi_Mode mode, gboolean use_continuation,
                        gboolean use_codegen){
    register Viterbi *viterbi = g_new0(Viterbi, 1);
    viterbi->vas = Viterbi_ArgumentSet_create(NULL);
    viterbi->name = Codegen_clean_path_component(name);
    if(use_continuation){ /* kept for API fidelity; never run (no reduced space) */
        viterbi->model = C4_Model_copy(model);
        C4_Model_configure_start_state(viterbi->model, C4_Scope_CORNER,
             viterbi->model->start_state->cell_start_func,
             viterbi->model->start_state->cell_start_macro);
        C4_Model_configure_end_state(viterbi->model, C4_Scope_CORNER,
             viterbi->model->end_state->cell_end_func,
             viterbi->model->end_state->cell_end_macro);
    } else {
        viterbi->model = C4_Model_share(model);
        }
    viterbi->func = NULL; /* there is no generated C: the tables ARE the compiled model */
    viterbi->mode = mode;
    viterbi->use_continuation = use_continuation;
    viterbi->cell_size = 1 + viterbi->model->total_shadow_designations;
    viterbi->layout = Layout_create(viterbi->model);
    return viterbi;
    }

void Viterbi_destroy(Viterbi *viterbi){
    register B200_Tables *bt, *prev = NULL;
    for(bt = tables_list; bt; prev = bt, bt = bt->next)
        if(bt->viterbi == viterbi){
            if(prev) prev->next = bt->next; else tables_list = bt->next;
            g_free(bt);
            break;
            }
    C4_Model_destroy(viterbi->model);
    Layout_destroy(viterbi->layout);
    g_free(viterbi->name);
    g_free(viterbi);
    return;
    }

/* The device keeps 4 bit/cell in a band; --dpmemory never forces the
 * region / checkpoint machinery (viterbi.c:128-150). */
gboolean Viterbi_use_reduced_space(Viterbi *viterbi, Region *region){
    return FALSE;
    }

Viterbi_Data *Viterbi_Data_create(Viterbi *viterbi, Region *region){
    register Viterbi_Data *vd = g_new0(Viterbi_Data, 1);
    if(viterbi->mode == Viterbi_Mode_FIND_REGION)
        vd->alignment_region = Region_copy(region);
    return vd;
    }

void Viterbi_Data_destroy(Viterbi_Data *vd){
    register B200_Path *path = (B200_Path*)vd->traceback;
    if(vd->alignment_region)
        Region_destroy(vd->alignment_region);
    if(path){
        g_free(path->ops);
        g_free(path);
        }
    g_free(vd);
    return;
    }

void Viterbi_Data_set_continuation(Viterbi_Data *vd,
              C4_State *first_state, C4_Score *first_cell,
              C4_State *final_state, C4_Score *final_cell){
    g_error("libc4b200: continuation DP is not used by the GPU engine");
    }

void Viterbi_Data_clear_continuation(Viterbi_Data *vd){
    return;
    }

GPtrArray *Viterbi_Checkpoint_traceback(Viterbi *viterbi,
        Viterbi_Data *vd, Region *region,
        C4_State *first_state, C4_Score *final_cell){
    g_error("libc4b200: checkpoint traceback is not used by the GPU engine");
    return NULL;
    }

void Viterbi_SubAlignment_destroy(Viterbi_SubAlignment *vsa){
    Region_destroy(vsa->region);
    g_free(vsa->final_cell);
    g_free(vsa);
    return;
    }

Codegen *Viterbi_make_Codegen(Viterbi *viterbi){
    g_error("libc4b200: nothing to generate for [%s]", viterbi->name);
    return NULL;
    }

/* ---- the call: Viterbi_calculate (viterbi.c:846-865) ----------------------------- */
void b200_fill_scoring(Match_ArgumentSet *mas, c4b_scoring *sc){
    register gint i, j;
    register Intron_ArgumentSet *ias = Intron_ArgumentSet_create(NULL);
    register Translate *t = mas->translate;
    memset(sc, 0, sizeof(c4b_scoring));
    for(i = 0; i < SUBMAT_ALPHABETSIZE; i++)
        for(j = 0; j < SUBMAT_ALPHABETSIZE; j++){
            sc->dna_matrix[i*C4B_SUBMAT_N+j] = mas->dna_submat->matrix[i][j];
            sc->protein_matrix[i*C4B_SUBMAT_N+j]
                = mas->protein_submat->matrix[i][j];
            }
    for(i = 0; i < 256; i++){
        sc->dna_index[i] = mas->dna_submat->index[i];
        sc->protein_index[i] = mas->protein_submat->index[i];
        sc->nt2d[i] = t->nt2d[i];
        }
    for(i = 0; i < 4096; i++)
        sc->codon_aa[i] = t->aa[t->trans[i]];
    sc->min_intron = ias->min_intron;
    sc->max_intron = ias->max_intron;
    return;
    }

gboolean b200_model_has_splice(c4b_model *m){
    register gint i;
    for(i = 0; i < m->n_calcs; i++)
        if((m->calcs[i].kind == C4B_CALC_SPLICE_PRE)
        || (m->calcs[i].kind == C4B_CALC_SPLICE_POST))
            return TRUE;
    return FALSE;
    }

gint b200_blocked_list(SubOpt *subopt, Region *region, gint32 **bq, gint32 **bt){
    register SubOpt_Index *soi = subopt?SubOpt_Index_create(subopt, region):NULL;
    register SubOpt_Index_Row *soir;
    register gint i, j, n_blocked = 0;
    (*bq) = (*bt) = NULL;
    if(!soi)
        return 0;
    /* rows are sorted by target_pos, positions by query_pos (subopt.c:250-338) */
    for(i = 0; i < soi->row_list->len; i++){
        soir = soi->row_list->pdata[i];
        if(soir != soi->blank_row)
            n_blocked += soir->total;
        }
    (*bq) = g_new(gint32, n_blocked+1);
    (*bt) = g_new(gint32, n_blocked+1);
    n_blocked = 0;
    for(i = 0; i < soi->row_list->len; i++){
        soir = soi->row_list->pdata[i];
        if(soir == soi->blank_row)
            continue;
        for(j = 0; j < soir->total; j++){
            (*bq)[n_blocked] = soir->query_pos[j];
            (*bt)[n_blocked++] = soir->target_pos;
            }
        }
    SubOpt_Index_destroy(soi);
    return n_blocked;
    }

/* what a Viterbi_DP_Func leaves in vd (viterbi.c:464-478,633-653) */
static void vd_set_result(Viterbi_Data *vd, Region *region, c4b_result *result,
                          B200_Path *path){
    vd->curr_query_end = result->query_end - region->query_start;
    vd->curr_target_end = result->target_end - region->target_start;
    vd->curr_query_start = result->query_start - region->query_start;
    vd->curr_target_start = result->target_start - region->target_start;
    if(vd->alignment_region){
        vd->alignment_region->query_start = result->query_start;
        vd->alignment_region->target_start = result->target_start;
        vd->alignment_region->query_length = result->query_end - result->query_start;
        vd->alignment_region->target_length = result->target_end - result->target_start;
        }
    if(path){
        path->result = (*result);
        if(vd->traceback){
            g_free(((B200_Path*)vd->traceback)->ops);
            g_free(vd->traceback);
            }
        vd->traceback = (C4_Transition****)path;
        }
    return;
    }

/* The batch hook (gam_b200.c) computed this call's answer ahead of time?  Only if it is
 * exactly the call the answer was computed for. */
static gboolean prefetch_lookup(Viterbi *viterbi, Region *region, Viterbi_Data *vd,
                                Ungapped_Data *ud, gint n_blocked, gint32 *bq, gint32 *bt,
                                C4_Score *score){
    register B200_Replay *rp = b200_replay;
    register B200_Round *round;
    register B200_Path *path;
    if(!rp || (rp->cursor >= rp->n_rounds))
        return FALSE;
    round = &rp->rounds[rp->cursor];
    if((rp->viterbi != viterbi) || (viterbi->mode != Viterbi_Mode_FIND_PATH)
    || (rp->query != ud->query) || (rp->target != ud->target)
    || region->query_start || region->target_start
    || (region->query_length != ud->query->len)
    || (region->target_length != ud->target->len)
    || (round->n_blocked != n_blocked)
    || (n_blocked && (memcmp(round->bq, bq, n_blocked*sizeof(gint32))
                   || memcmp(round->bt, bt, n_blocked*sizeof(gint32))))){
        rp->cursor = rp->n_rounds; /* the series diverged: nothing later can match */
        b200_stat_prefetch_misses++;
        return FALSE;
        }
    rp->cursor++;
    path = g_new0(B200_Path, 1);
    path->ops = g_new(gint32, 2*round->result.n_ops+2);
    memcpy(path->ops, round->ops, 2*round->result.n_ops*sizeof(gint32));
    vd_set_result(vd, region, &round->result, path);
    path->result.ops_offset = 0;
    (*score) = round->result.score;
    b200_stat_prefetch_hits++;
    return TRUE;
    }

/* ---- BSDP prefetch: kept FIND_SCORE answers of the current comparison ------------------ */
typedef struct B200_ScoreEntry {
    Viterbi *viterbi;
    gint qs, ts, ql, tl;
    gint n_blocked;
    gint32 *bq, *bt;
    c4b_result result;
    struct B200_ScoreEntry *next;
} B200_ScoreEntry;
#define SCORE_BUCKETS 4096
static B200_ScoreEntry *score_table[SCORE_BUCKETS];
static gint score_entries = 0;

static guint score_hash(Viterbi *viterbi, Region *region){
    register guint64 h = (guint64)(gsize)viterbi * 0x9E3779B97F4A7C15ull;
    h ^= ((guint64)(guint)region->query_start << 32) | (guint)region->target_start;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= ((guint64)(guint)region->query_length << 32) | (guint)region->target_length;
    h *= 0xC4CEB9FE1A85EC53ull;
    return (guint)(h >> 40) & (SCORE_BUCKETS-1);
    }

static void score_cache_clear(void){
    register gint i;
    register B200_ScoreEntry *e, *next;
    if(!score_entries)
        return;
    for(i = 0; i < SCORE_BUCKETS; i++){
        for(e = score_table[i]; e; e = next){
            next = e->next;
            g_free(e->bq);
            g_free(e->bt);
            g_free(e);
            }
        score_table[i] = NULL;
        }
    score_entries = 0;
    return;
    }

static B200_ScoreEntry *score_cache_find(Viterbi *viterbi, Region *region,
                                         gint n_blocked, gint32 *bq, gint32 *bt){
    register B200_ScoreEntry *e;
    if(!score_entries)
        return NULL;
    for(e = score_table[score_hash(viterbi, region)]; e; e = e->next)
        if((e->viterbi == viterbi) && (e->qs == region->query_start)
        && (e->ts == region->target_start) && (e->ql == region->query_length)
        && (e->tl == region->target_length) && (e->n_blocked == n_blocked)
        && (!n_blocked || (!memcmp(e->bq, bq, n_blocked*sizeof(gint32))
                        && !memcmp(e->bt, bt, n_blocked*sizeof(gint32)))))
            return e;
    return NULL;
    }

void b200_prefetch_scores(Viterbi *viterbi, gint n, Region **regions, gpointer user_data,
                          SubOpt *subopt){
    register Ungapped_Data *ud = user_data;
    register c4b_model *tables = b200_tables_for(viterbi);
    register gboolean with_splice = b200_model_has_splice(tables);
    register c4b_pair *pairs;
    register c4b_result *results;
    register B200_ScoreEntry **entry;
    register gint i, k, m = 0;
    register guint h;
    gint32 *bq, *bt;
    gint nb;
    c4b_batch *batch = NULL;
    c4b_scoring scoring;
    register gdouble t_begin = now_seconds();
    if((n <= 0) || (viterbi->mode != Viterbi_Mode_FIND_SCORE) || model_is_bound(viterbi->model)
    || viterbi->model->start_state->cell_start_func || viterbi->model->end_state->cell_end_func)
        return;
    pair_cache_fetch(ud, with_splice);
    b200_fill_scoring(ud->mas, &scoring);
    pairs = g_new0(c4b_pair, n);
    results = g_new0(c4b_result, n);
    entry = g_new0(B200_ScoreEntry*, n);
    for(i = 0; i < n; i++){
        nb = b200_blocked_list((subopt && subopt->path_count)?subopt:NULL, regions[i], &bq, &bt);
        if(score_cache_find(viterbi, regions[i], nb, bq, bt)){ /* asked for twice */
            g_free(bq);
            g_free(bt);
            continue;
            }
        entry[m] = g_new0(B200_ScoreEntry, 1);
        entry[m]->viterbi = viterbi;
        entry[m]->qs = regions[i]->query_start;
        entry[m]->ts = regions[i]->target_start;
        entry[m]->ql = regions[i]->query_length;
        entry[m]->tl = regions[i]->target_length;
        entry[m]->n_blocked = nb;
        entry[m]->bq = bq;
        entry[m]->bt = bt;
        h = score_hash(viterbi, regions[i]); /* visible to the duplicate test of later regions */
        entry[m]->next = score_table[h];
        score_table[h] = entry[m];
        score_entries++;
        pairs[m].query = (const uint8_t*)pair_cache.qseq;
        pairs[m].target = (const uint8_t*)pair_cache.tseq;
        pairs[m].query_len = ud->query->len;
        pairs[m].target_len = ud->target->len;
        pairs[m].query_start = regions[i]->query_start;
        pairs[m].target_start = regions[i]->target_start;
        pairs[m].query_length = regions[i]->query_length;
        pairs[m].target_length = regions[i]->target_length;
        pairs[m].blocked_query_pos = bq;
        pairs[m].blocked_target_pos = bt;
        pairs[m].n_blocked = nb;
        pairs[m].reserved = C4B_PAIR_BUFFERS_STABLE;
        if(with_splice)
            for(k = 0; k < 4; k++)
                pairs[m].splice[k] = pair_cache.splice + (gsize)k*ud->target->len;
        m++;
        }
    if(m){
        if(c4b_batch_create(get_engine(), tables, &scoring, m, pairs, 0, &batch)
        || c4b_batch_run(batch, C4B_IMPOSSIBLY_LOW_SCORE)
        || c4b_batch_fetch(batch, results, NULL, 0))
            g_error("libc4b200: %s", c4b_last_error());
        c4b_batch_destroy(batch);
        for(i = 0; i < m; i++)
            entry[i]->result = results[i];
        b200_stat_score_prefetched += m;
        b200_stat_score_batches++;
        }
    g_free(entry);
    g_free(results);
    g_free(pairs);
    stat_prefetch += now_seconds() - t_begin;
    return;
    }

static void fill_pair_region(c4b_pair *pair, Ungapped_Data *ud, Region *region, gboolean with_splice,
                             SubOpt *subopt, gint32 **bq, gint32 **bt){
    register gint k;
    memset(pair, 0, sizeof(c4b_pair));
    pair->query = (const uint8_t*)pair_cache.qseq;
    pair->target = (const uint8_t*)pair_cache.tseq;
    pair->query_len = ud->query->len;
    pair->target_len = ud->target->len;
    pair->query_start = region->query_start;
    pair->target_start = region->target_start;
    pair->query_length = region->query_length;
    pair->target_length = region->target_length;
    pair->reserved = C4B_PAIR_BUFFERS_STABLE;
    pair->n_blocked = b200_blocked_list((subopt && subopt->path_count)?subopt:NULL, region, bq, bt);
    pair->blocked_query_pos = (*bq);
    pair->blocked_target_pos = (*bt);
    if(with_splice)
        for(k = 0; k < 4; k++)
            pair->splice[k] = pair_cache.splice + (gsize)k*ud->target->len;
    return;
    }

void b200_span_scores(gpointer heuristic_span, gint n, Region **src_regions, Region **dst_regions,
                      gpointer user_data, SubOpt *subopt, C4_Score *scores){
    register Heuristic_Span *hs = heuristic_span;
    register Ungapped_Data *ud = user_data;
    register c4b_model *src_tables = b200_tables_for(hs->src_optimal->find_score),
                       *dst_tables = b200_tables_for(hs->dst_optimal->find_score);
    register gboolean with_splice = b200_model_has_splice(src_tables)
                                 || b200_model_has_splice(dst_tables);
    register c4b_span_job *jobs = g_new0(c4b_span_job, n);
    register gint32 **lists = g_new0(gint32*, 4*n);
    register gint i;
    register gdouble t_begin = now_seconds();
    c4b_scoring scoring;
    pair_cache_fetch(ud, with_splice);
    b200_fill_scoring(ud->mas, &scoring);
    for(i = 0; i < n; i++){
        fill_pair_region(&jobs[i].src, ud, src_regions[i], with_splice, subopt,
                         &lists[4*i], &lists[4*i+1]);
        fill_pair_region(&jobs[i].dst, ud, dst_regions[i], with_splice, subopt,
                         &lists[4*i+2], &lists[4*i+3]);
        jobs[i].span[0] = hs->span->min_query;
        jobs[i].span[1] = hs->span->max_query;
        jobs[i].span[2] = hs->span->min_target;
        jobs[i].span[3] = hs->span->max_target;
        }
    if(c4b_span_score_batch(get_engine(), src_tables, dst_tables, &scoring, n, jobs, scores))
        g_error("libc4b200: %s", c4b_last_error());
    for(i = 0; i < 4*n; i++)
        g_free(lists[i]);
    g_free(lists);
    g_free(jobs);
    b200_stat_score_prefetched += 2*n;
    b200_stat_score_batches++;
    stat_spans += now_seconds() - t_begin;
    return;
    }

C4_Score Viterbi_calculate(Viterbi *viterbi, Region *region,
                           Viterbi_Data *vd, gpointer user_data,
                           SubOpt *subopt){
    register Ungapped_Data *ud = user_data; /* every shipped *_Data inherits it */
    register c4b_model *tables = b200_tables_for(viterbi);
    register gchar *qseq, *tseq;
    register gint i, j, n_blocked = 0, mode = 0;
    gint32 *bq = NULL, *bt = NULL;
    C4_Score prefetched_score;
    register gint64 ops_capacity;
    register B200_Path *path = NULL;
    register gdouble t_begin = now_seconds(), t_prepared, t_done;
    static gboolean stats_registered = FALSE;
    c4b_scoring scoring;
    c4b_pair pair;
    c4b_result result;
    g_assert(Region_is_valid(region));
    if(vd->continuation || (viterbi->mode == Viterbi_Mode_FIND_CHECKPOINTS))
        g_error("libc4b200: continuation / checkpoint modes are not used");
    if(model_is_bound(viterbi->model)){
        /* user_data is the bound matrix [query_range+1][target_range+1]; the model reads
         * no symbol (all calcs are constants), so the sequences are placeholders */
        register C4_Score **matrix = user_data;
        register gint Q = region->query_length, T = region->target_length;
        register c4b_score *flat = g_new(c4b_score, (gsize)(Q+1)*(T+1));
        qseq = g_strnfill(region->query_start+Q+4, 'A');
        tseq = g_strnfill(region->target_start+T+4, 'A');
        memset(&pair, 0, sizeof(pair));
        memset(&scoring, 0, sizeof(scoring));
        pair.query = (const uint8_t*)qseq;
        pair.target = (const uint8_t*)tseq;
        pair.query_len = region->query_start+Q;
        pair.target_len = region->target_start+T;
        pair.query_start = region->query_start;
        pair.target_start = region->target_start;
        pair.query_length = Q;
        pair.target_length = T;
        for(i = 0; i <= Q; i++)
            for(j = 0; j <= T; j++)
                flat[(gsize)i*(T+1)+j] = matrix[region->query_start+i][region->target_start+j];
        if(c4b_viterbi_calculate_cells(get_engine(), tables, &scoring, &pair, 0,
                                       NULL, flat, &result, NULL, 0))
            g_error("libc4b200: %s", c4b_last_error());
        for(i = 0; i <= Q; i++)
            for(j = 0; j <= T; j++)
                matrix[region->query_start+i][region->target_start+j] = flat[(gsize)i*(T+1)+j];
        g_free(flat);
        g_free(qseq);
        g_free(tseq);
        vd->curr_query_end = result.query_end - region->query_start;
        vd->curr_target_end = result.target_end - region->target_start;
        if(g_getenv("EXONERATE_B200_TRACE")){
            register gint64 sum = 0;
            for(i = 0; i <= Q; i++)
                for(j = 0; j <= T; j++)
                    sum += matrix[region->query_start+i][region->target_start+j];
            fprintf(stderr, "b200-trace bound matrix sum %ld corner %d\n", (glong)sum,
                    matrix[region->query_start+Q][region->target_start+T]);
            }
        trace_call(viterbi, region, "bound", result.score);
        return result.score;
        }
    switch(viterbi->mode){
        case Viterbi_Mode_FIND_SCORE:  mode = 0; break;
        case Viterbi_Mode_FIND_PATH:   mode = 1; break;
        case Viterbi_Mode_FIND_REGION: mode = 2; break;
        default: g_error("libc4b200: bad mode"); break;
        }
    if(!stats_registered){
        stats_registered = TRUE;
        if(g_getenv("EXONERATE_B200_STATS"))
            atexit(print_viterbi_stats);
        }
    n_blocked = b200_blocked_list(subopt, region, &bq, &bt);
    if(b200_replay && prefetch_lookup(viterbi, region, vd, ud, n_blocked, bq, bt,
                                      &prefetched_score)){
        g_free(bq);
        g_free(bt);
        stat_calls++;
        stat_total += now_seconds() - t_begin;
        trace_call(viterbi, region, "replay", prefetched_score);
        return prefetched_score;
        }
    if((mode == 0) && score_entries && (pair_cache.query == ud->query)
    && (pair_cache.target == ud->target)){
        register B200_ScoreEntry *se = score_cache_find(viterbi, region, n_blocked, bq, bt);
        if(se){ /* this very fill was part of the comparison's prefetched batch */
            vd_set_result(vd, region, &se->result, NULL);
            g_free(bq);
            g_free(bt);
            b200_stat_score_hits++;
            stat_calls++;
            stat_total += now_seconds() - t_begin;
            trace_call(viterbi, region, "prefetched", se->result.score);
            return se->result.score;
            }
        }
    pair_cache_fetch(ud, b200_model_has_splice(tables));
    qseq = pair_cache.qseq;
    tseq = pair_cache.tseq;
    memset(&pair, 0, sizeof(pair));
    pair.query = (const uint8_t*)qseq;
    pair.target = (const uint8_t*)tseq;
    pair.query_len = ud->query->len;
    pair.target_len = ud->target->len;
    pair.query_start = region->query_start;
    pair.target_start = region->target_start;
    pair.query_length = region->query_length;
    pair.target_length = region->target_length;
    pair.reserved = C4B_PAIR_BUFFERS_STABLE; /* pair_cache owns them until the next comparison */
    b200_fill_scoring(ud->mas, &scoring);
    if(b200_model_has_splice(tables))
        for(i = 0; i < 4; i++)
            pair.splice[i] = pair_cache.splice + (gsize)i*ud->target->len;
    if(n_blocked){
        pair.blocked_query_pos = bq;
        pair.blocked_target_pos = bt;
        pair.n_blocked = n_blocked;
        }
    t_prepared = now_seconds();
    ops_capacity = (gint64)region->query_length + region->target_length + 8;
    if(mode == 1){
        path = g_new0(B200_Path, 1);
        path->ops = g_new(gint32, 2*ops_capacity);
        }
    if(viterbi->model->start_state->cell_start_func
    || viterbi->model->end_state->cell_end_func){
        /* BSDP span models (src/bsdp/heuristic.c:445-528): the callbacks read / write the
         * integration matrices on the host; the device gets START's cell of every lattice
         * cell as a table and returns END's cell wherever END was reached */
        register gint C = 1 + viterbi->model->total_shadow_designations,
                      Q = region->query_length, T = region->target_length;
        register gsize ncell = (gsize)(Q+1)*(T+1);
        register c4b_score *start_cells = NULL, *end_cells = NULL, *cell;
        register gint l;
        if(mode == 2)
            g_error("libc4b200: FIND_REGION on a model with cell callbacks");
        if(viterbi->model->start_state->cell_start_func){
            start_cells = g_new(c4b_score, ncell*C);
            for(i = 0; i <= Q; i++)
                for(j = 0; j <= T; j++){
                    cell = viterbi->model->start_state->cell_start_func(
                               region->query_start+i, region->target_start+j, user_data);
                    for(l = 0; l < C; l++)
                        start_cells[((gsize)i*(T+1)+j)*C+l] = cell[l];
                    }
            }
        if(viterbi->model->end_state->cell_end_func){
            end_cells = g_new(c4b_score, ncell*C);
            for(i = 0; i < ncell*C; i++)
                end_cells[i] = C4_IMPOSSIBLY_LOW_SCORE - 1; /* "END not reached" */
            }
        if(c4b_viterbi_calculate_cells(get_engine(), tables, &scoring, &pair, mode,
                start_cells, end_cells, &result, path?path->ops:NULL, ops_capacity))
            g_error("libc4b200: %s", c4b_last_error());
        if(end_cells){ /* same visiting order as the fill: target outer, query inner */
            for(j = 0; j <= T; j++)
                for(i = 0; i <= Q; i++){
                    cell = &end_cells[((gsize)i*(T+1)+j)*C];
                    if(cell[0] != (C4_IMPOSSIBLY_LOW_SCORE - 1))
                        viterbi->model->end_state->cell_end_func(cell, C,
                            region->query_start+i, region->target_start+j, user_data);
                    }
            g_free(end_cells);
            }
        if(start_cells)
            g_free(start_cells);
    } else if(c4b_viterbi_calculate(get_engine(), tables, &scoring, &pair, mode,
                             &result, path?path->ops:NULL, ops_capacity))
        g_error("libc4b200: %s", c4b_last_error());
    stat_engine += now_seconds() - t_prepared;
    vd_set_result(vd, region, &result, path);
    g_free(bq);
    g_free(bt);
    t_done = now_seconds();
    stat_calls++;
    stat_total += t_done - t_begin;
    stat_prepare += t_prepared - t_begin;
    trace_call(viterbi, region, "device", result.score);
    return result.score;
    }

/* Viterbi_Data_create_Alignment (viterbi.c:342-392): the walk already happened
 * on the device; replay the (transition, length) list into an Alignment. */
Alignment *Viterbi_Data_create_Alignment(Viterbi_Data *vd,
                  C4_Model *model, C4_Score score, Region *region){
    register B200_Path *path = (B200_Path*)vd->traceback;
    register Alignment *alignment;
    register Region *alignment_region;
    register gint i;
    g_assert(path);
    alignment_region = Region_create(path->result.query_start,
        path->result.target_start,
        path->result.query_end - path->result.query_start,
        path->result.target_end - path->result.target_start);
    alignment = Alignment_create(model, alignment_region, score);
    for(i = 0; i < path->result.n_ops; i++)
        Alignment_add(alignment,
            model->transition_list->pdata[path->ops[2*(path->result.ops_offset+i)]],
            path->ops[2*(path->result.ops_offset+i)+1]);
    Region_destroy(alignment_region);
    return alignment;
    }
