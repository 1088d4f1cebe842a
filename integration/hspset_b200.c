/* hspset_b200.c -- HSP seeding of the unmodified reference on the device.
 *
 * The reference's hspset.o stays in the link, with two of its entry points renamed by
 * objcopy (integration/Makefile):
 *     HSPset_seed_hsp  -> c4bref_HSPset_seed_hsp      (src/comparison/hspset.c:933-997)
 *     HSPset_finalise  -> c4bref_HSPset_finalise      (src/comparison/hspset.c:1123-1150)
 * (and Comparison_has_hsps in comparison.o, src/comparison/comparison.c:191-202)
 * and this object provides them instead.  The seeder streams one HSPset_seed_hsp per word
 * hit (src/comparison/seeder.c:645) and then calls HSPset_finalise: we only COLLECT the
 * seeds, and at finalise extend all of them in one c4b_hsp_extend_batch call, replay the
 * diagonal horizon over the results in arrival order, and hand every surviving HSP to the
 * reference's own HSPset_add_known_hsp (HSP_init + HSP_store: threshold, --hspfilter
 * queues, hsp_list) -- so everything downstream of the extension is still the reference.
 * Every Match_Type (DNA2DNA, PROTEIN2PROTEIN, DNA2PROTEIN, PROTEIN2DNA, CODON2CODON) has a
 * device form; an unknown one is a g_error (there is no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "hspset.h"
#include "comparison.h"
#include "match.h"
#include "sequence.h"
#include "submat.h"
#include "translate.h"
#include "c4b200.h"

extern void c4bref_HSPset_seed_hsp(HSPset *hsp_set, guint query_start, guint target_start);
extern HSPset *c4bref_HSPset_finalise(HSPset *hsp_set);
extern gboolean c4bref_Comparison_has_hsps(Comparison *comparison);
extern c4b_engine *exonerate_b200_engine(void); /* viterbi_b200.c */

typedef struct B200_Pending {
    HSPset *hsp_set;
    guint *seeds; /* (query_start, target_start) pairs in arrival order */
    gint n, cap;
    struct B200_Pending *next;
} B200_Pending;
static B200_Pending *pending_list = NULL;

/* EXONERATE_B200_STATS=1: one line on stderr at exit (the CLI tests check the device was used) */
static glong stat_batches = 0, stat_seeds = 0, stat_passthrough = 0;
static gdouble stat_flatten = 0, stat_device = 0;
static gdouble now_seconds(void){
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9*ts.tv_nsec;
    }
static void print_stats(void){
    fprintf(stderr, "exonerate_b200: hsp batches %ld, seeds extended on the device %ld, "
                    "seeds passed to the reference %ld (flatten + mask %.3f s, device call %.3f s)\n",
            stat_batches, stat_seeds, stat_passthrough, stat_flatten, stat_device);
    }
static void count(glong *what, glong by){
    static gboolean registered = FALSE;
    if(!registered){
        registered = TRUE;
        if(g_getenv("EXONERATE_B200_STATS"))
            atexit(print_stats);
        }
    (*what) += by;
    }

static gint device_match_kind(Match *match){
    switch(match->type){
        case Match_Type_DNA2DNA:         return C4B_CALC_MATCH_DNA;
        case Match_Type_PROTEIN2PROTEIN: return C4B_CALC_MATCH_PROTEIN;
        case Match_Type_PROTEIN2DNA:     return C4B_CALC_MATCH_1_3;
        case Match_Type_DNA2PROTEIN:     return C4B_CALC_MATCH_3_1;
        case Match_Type_CODON2CODON:     return C4B_CALC_MATCH_3_3;
        default:                         return -1;
        }
    }

void HSPset_seed_hsp(HSPset *hsp_set, guint query_start, guint target_start){
    register B200_Pending *p;
    if(device_match_kind(hsp_set->param->match) < 0) /* no CPU fallback: hard error */
        g_error("libc4b200: Match_Type [%d] has no device form for HSP extension",
                hsp_set->param->match->type);
    for(p = pending_list; p; p = p->next)
        if(p->hsp_set == hsp_set)
            break;
    if(!p){
        p = g_new0(B200_Pending, 1);
        p->hsp_set = hsp_set;
        p->cap = 1024;
        p->seeds = g_new(guint, p->cap << 1);
        p->next = pending_list;
        pending_list = p;
        }
    if(p->n == p->cap){
        p->cap <<= 1;
        p->seeds = g_renew(guint, p->seeds, p->cap << 1);
        }
    p->seeds[p->n << 1] = query_start;
    p->seeds[(p->n << 1) + 1] = target_start;
    p->n++;
    return;
    }

static guint8 *mask_bytes(Sequence *seq, gchar *flat){
    register guint8 *mask = g_new(guint8, seq->len + 4);
    register gint i;
    register gboolean any = FALSE;
    for(i = 0; i < seq->len; i++){
        mask[i] = Alphabet_is_masked(seq->alphabet, (guchar)flat[i]) ? 1 : 0;
        any |= mask[i];
        }
    if(!any){
        g_free(mask);
        return NULL;
        }
    return mask;
    }

static void flush(B200_Pending *p){
    register HSPset *hsp_set = p->hsp_set;
    register HSP_Param *param = hsp_set->param;
    register Match *match = param->match;
    register Match_ArgumentSet *mas = match->mas;
    register Translate *tr = mas->translate;
    register gint qadv = match->query->advance, tadv = match->target->advance;
    register gint i, j, k, diag_pos, section_pos, query_frame, target_frame;
    register gchar *qflat = g_new(gchar, hsp_set->query->len + 4),
                   *tflat = g_new(gchar, hsp_set->target->len + 4);
    register guint8 *qmask, *tmask;
    register c4b_scoring *scoring = g_new0(c4b_scoring, 1);
    register c4b_hsp *ext = g_new(c4b_hsp, p->n);
    c4b_hsp_param hp;
    register gdouble t0 = now_seconds(), t1;
    Sequence_strncpy(hsp_set->query, 0, hsp_set->query->len, qflat);
    Sequence_strncpy(hsp_set->target, 0, hsp_set->target->len, tflat);
    qmask = mask_bytes(hsp_set->query, qflat);
    tmask = mask_bytes(hsp_set->target, tflat);
    for(i = 0; i < SUBMAT_ALPHABETSIZE; i++)
        for(j = 0; j < SUBMAT_ALPHABETSIZE; j++){
            scoring->dna_matrix[i*C4B_SUBMAT_N+j] = mas->dna_submat->matrix[i][j];
            scoring->protein_matrix[i*C4B_SUBMAT_N+j] = mas->protein_submat->matrix[i][j];
            }
    for(i = 0; i < 256; i++){
        scoring->dna_index[i] = mas->dna_submat->index[i];
        scoring->protein_index[i] = mas->protein_submat->index[i];
        scoring->nt2d[i] = tr->nt2d[i];
        }
    for(i = 0; i < 4096; i++)
        scoring->codon_aa[i] = tr->aa[tr->trans[i]];
    hp.match_kind = device_match_kind(match);
    hp.seedlen = param->seedlen;
    hp.dropoff = param->dropoff;
    hp.threshold = param->threshold;
    t1 = now_seconds();
    stat_flatten += t1 - t0;
    if(c4b_hsp_extend_batch(exonerate_b200_engine(), scoring, &hp,
            (const uint8_t*)qflat, hsp_set->query->len, qmask,
            (const uint8_t*)tflat, hsp_set->target->len, tmask,
            p->n, (const c4b_hsp_seed*)p->seeds, ext))
        g_error("libc4b200: %s", c4b_last_error());
    stat_device += now_seconds() - t1;
    count(&stat_batches, 1);
    count(&stat_seeds, p->n);
    /* the diagonal horizon, as HSPset_seed_hsp keeps it (hspset.c:935-972,991-996):
     * [0] target end of the last HSP on the diagonal section, [1] seeds seen since,
     * [2] which diagonal the section currently holds */
    for(k = 0; k < p->n; k++){
        register guint query_start = p->seeds[k << 1], target_start = p->seeds[(k << 1) + 1];
        diag_pos = (target_start * qadv) - (query_start * tadv);
        query_frame = query_start % qadv;
        target_frame = target_start % tadv;
        section_pos = (diag_pos + hsp_set->query->len) % hsp_set->query->len;
        if(param->seed_repeat > 1){
            if(hsp_set->horizon[2][section_pos][query_frame][target_frame]
               != (diag_pos + hsp_set->query->len)){
                hsp_set->horizon[0][section_pos][query_frame][target_frame] = 0;
                hsp_set->horizon[1][section_pos][query_frame][target_frame] = 0;
                hsp_set->horizon[2][section_pos][query_frame][target_frame]
                    = diag_pos + hsp_set->query->len;
                }
            }
        if(target_start < hsp_set->horizon[0][section_pos][query_frame][target_frame])
            continue;
        if(param->seed_repeat > 1){
            if(++hsp_set->horizon[1][section_pos][query_frame][target_frame] < param->seed_repeat)
                continue;
            hsp_set->horizon[1][section_pos][query_frame][target_frame] = 0;
            }
        if(ext[k].status)
            g_error("Initial HSP score [%d] less than zero", ext[k].score);
        if(ext[k].stored)
            HSPset_add_known_hsp(hsp_set, ext[k].query_start, ext[k].target_start, ext[k].length);
        hsp_set->horizon[0][section_pos][query_frame][target_frame] = ext[k].target_end;
        }
    g_free(ext);
    g_free(scoring);
    if(qmask) g_free(qmask);
    if(tmask) g_free(tmask);
    g_free(qflat);
    g_free(tflat);
    return;
    }

static void flush_pending(HSPset *hsp_set){
    register B200_Pending *p, *prev = NULL;
    if(!hsp_set)
        return;
    for(p = pending_list; p; prev = p, p = p->next)
        if(p->hsp_set == hsp_set){
            if(prev) prev->next = p->next; else pending_list = p->next;
            if(p->n)
                flush(p);
            g_free(p->seeds);
            g_free(p);
            break;
            }
    return;
    }

HSPset *HSPset_finalise(HSPset *hsp_set){
    flush_pending(hsp_set);
    return c4bref_HSPset_finalise(hsp_set);
    }

/* The seeder asks Comparison_has_hsps() BEFORE it finalises (src/comparison/seeder.c:
 * 905-906, HSPset_is_empty reads hsp_set->is_empty): the collected seeds have to be
 * extended by then.  comparison.o is linked with this symbol renamed as well. */
gboolean Comparison_has_hsps(Comparison *comparison){
    flush_pending(comparison->dna_hspset);
    flush_pending(comparison->protein_hspset);
    flush_pending(comparison->codon_hspset);
    return c4bref_Comparison_has_hsps(comparison);
    }
