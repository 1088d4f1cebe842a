/* bsdp_b200.c -- the region fills of one heuristic (BSDP) comparison as device batches.
 *
 * BSDP scores a comparison lazily: HPair_finalise builds a graph with one node per HSP and one
 * edge per joinable HSP pair, every node / edge carrying a "sub-alignment region" (SAR,
 * src/bsdp/sar.c) scored first by a cheap bound; the search then CONFIRMS bounds on demand, one
 * Optimal_find_score call on a small region each (SAR_Terminal_find_score sar.c:393-398,
 * SAR_Join_find_score :697-700; callers src/bsdp/hpair.c:147-290).  One synchronous device call
 * per tiny lattice is latency, not throughput (profiles/r01d_bsdp_cli.md: 0.5 ms per call).
 *
 * hpair.o is linked as a copy in which three external references are renamed by objcopy
 * (integration/Makefile):
 *     SAR_Terminal_create -> b200_SAR_Terminal_create      (sar.c:317-371)
 *     SAR_Join_create     -> b200_SAR_Join_create          (sar.c:579-675)
 *     BSDP_initialise     -> b200_BSDP_initialise          (called last by HPair_finalise, hpair.c:665-674)
 * The create wrappers call the reference's own functions and note {Optimal, Region} of every SAR
 * they return; b200_BSDP_initialise -- the moment the graph is complete and before the search
 * asks for the first score -- runs ALL noted terminal and join fills of the comparison, one
 * batch per derived model (b200_prefetch_scores), and then calls the reference's BSDP_initialise.
 * The search is untouched: its SAR_*_find_score -> Optimal_find_score -> Viterbi_calculate calls
 * find the kept answers (exact match on Viterbi, region and SubOpt blocked cells; a region a
 * reported alignment has since touched simply misses and runs synchronously).
 *
 * Span edges (SAR_Span_find_score, sar.c:898-917: a src fill that reports END's cell of every
 * cell, Heuristic_Span_integrate, a dst fill that starts from the integrated cells) go the same
 * way with two more renamed references, SAR_Span_create (sar.c:706-854) and SAR_Span_find_score:
 * every span edge the graph keeps is noted, all edges of one Heuristic_Span are scored by ONE
 * c4b_span_score_batch call (both fills, the integration and the START tables on the device),
 * and b200_SAR_Span_find_score answers from the kept score when the SubOpt blocked cells of both
 * regions are what they were at that moment -- else it calls the reference's function.
 *
 * EXONERATE_B200_BSDP_BATCH=0 switches the prefetch off. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hpair.h"
#include "sar.h"
#include "bsdp.h"
#include "b200_binding.h"

typedef struct {
    HPair *hpair;
    Optimal *optimal;
    Region *region; /* owned by the SAR, which outlives b200_BSDP_initialise */
} B200_Noted;

static B200_Noted *noted = NULL;
static gint noted_n = 0, noted_cap = 0;

static gboolean enabled(void){
    static gint on = -1;
    if(on < 0){
        register const gchar *v = g_getenv("EXONERATE_B200_BSDP_BATCH");
        on = (v && !atoi(v))?0:1;
        }
    return on;
    }

static void note(HPair *hpair, Optimal *optimal, Region *region){
    if(!enabled() || !optimal || !optimal->find_score)
        return;
    if(noted_n == noted_cap){
        noted_cap = noted_cap?noted_cap*2:256;
        noted = g_renew(B200_Noted, noted, noted_cap);
        }
    noted[noted_n].hpair = hpair;
    noted[noted_n].optimal = optimal;
    noted[noted_n].region = region;
    noted_n++;
    return;
    }

SAR_Terminal *b200_SAR_Terminal_create(HSP *hsp, HPair *hpair, Heuristic_Match *match,
                                       gboolean is_start){
    register SAR_Terminal *sar_terminal = SAR_Terminal_create(hsp, hpair, match, is_start);
    if(sar_terminal)
        note(hpair, is_start?match->start_terminal->optimal:match->end_terminal->optimal,
             sar_terminal->region);
    return sar_terminal;
    }

SAR_Join *b200_SAR_Join_create(HSP *src_hsp, HSP *dst_hsp, HPair *hpair, Heuristic_Pair *pair){
    register SAR_Join *sar_join = SAR_Join_create(src_hsp, dst_hsp, hpair, pair);
    if(sar_join)
        note(hpair, pair->join->optimal, sar_join->region);
    return sar_join;
    }

/* ---- span edges ------------------------------------------------------------------------------- */
typedef struct {
    HPair *hpair;
    SAR_Span *sar_span; /* owned by the BSDP edge (or destroyed right after creation: see below) */
    Heuristic_Span *span;
    Region src, dst;    /* by value: a destroyed SAR_Span must not be mistaken for a new one */
    gboolean scored;
    C4_Score score;
    gint nb[2];
    gint32 *bq[2], *bt[2];
} B200_SpanNote;

static B200_SpanNote *span_note = NULL;
static gint span_n = 0, span_cap = 0;
static glong stat_span_hits = 0, stat_span_misses = 0;

static void span_notes_clear(void){
    register gint i, k;
    for(i = 0; i < span_n; i++)
        for(k = 0; k < 2; k++){
            g_free(span_note[i].bq[k]);
            g_free(span_note[i].bt[k]);
            }
    span_n = 0;
    return;
    }

static void print_span_stats(void){
    fprintf(stderr, "exonerate_b200: BSDP span edges answered from a batch %ld, scored per call %ld\n",
            stat_span_hits, stat_span_misses);
    }

SAR_Span *b200_SAR_Span_create(HSP *src_hsp, HSP *dst_hsp, HPair *hpair, Heuristic_Span *span,
                               C4_Portal *src_portal, C4_Portal *dst_portal){
    register SAR_Span *sar_span = SAR_Span_create(src_hsp, dst_hsp, hpair, span, src_portal,
                                                  dst_portal);
    register B200_SpanNote *sn;
    if(!sar_span || !enabled())
        return sar_span;
    if(span_n == span_cap){
        span_cap = span_cap?span_cap*2:256;
        span_note = g_renew(B200_SpanNote, span_note, span_cap);
        }
    sn = &span_note[span_n++];
    memset(sn, 0, sizeof(B200_SpanNote));
    sn->hpair = hpair;
    sn->sar_span = sar_span;
    sn->span = span;
    sn->src = (*sar_span->src_region);
    sn->dst = (*sar_span->dst_region);
    return sar_span;
    }

static gboolean same_region(Region *a, Region *b){
    return (a->query_start == b->query_start) && (a->target_start == b->target_start)
        && (a->query_length == b->query_length) && (a->target_length == b->target_length);
    }

C4_Score b200_SAR_Span_find_score(SAR_Span *sar_span, HPair *hpair){
    register gint i, k, nb;
    register B200_SpanNote *sn;
    register gboolean same;
    static gboolean registered = FALSE;
    gint32 *bq, *bt;
    if(!registered){
        registered = TRUE;
        if(g_getenv("EXONERATE_B200_STATS"))
            atexit(print_span_stats);
        }
    for(i = 0; i < span_n; i++){
        sn = &span_note[i];
        if(!sn->scored || (sn->sar_span != sar_span) || (sn->hpair != hpair)
        || (sn->span != sar_span->span) || !same_region(&sn->src, sar_span->src_region)
        || !same_region(&sn->dst, sar_span->dst_region))
            continue;
        same = TRUE; /* SubOpt blocked cells of both regions as they were when the batch ran? */
        for(k = 0; (k < 2) && same; k++){
            nb = b200_blocked_list((hpair->subopt && hpair->subopt->path_count)?hpair->subopt:NULL,
                                   k?sar_span->dst_region:sar_span->src_region, &bq, &bt);
            same = (nb == sn->nb[k])
                && (!nb || (!memcmp(bq, sn->bq[k], nb*sizeof(gint32))
                         && !memcmp(bt, sn->bt[k], nb*sizeof(gint32))));
            g_free(bq);
            g_free(bt);
            }
        if(!same)
            break;
        stat_span_hits++;
        return sn->score - (sar_span->src_component + sar_span->dst_component);
        }
    stat_span_misses++;
    return SAR_Span_find_score(sar_span, hpair);
    }

static void prefetch_spans(HPair *hpair){
    register gint i, j, k, n;
    register Region **src = g_new(Region*, span_n+1), **dst = g_new(Region*, span_n+1);
    register gint *which = g_new(gint, span_n+1);
    register C4_Score *scores = g_new(C4_Score, span_n+1);
    register Heuristic_Span *span;
    register SubOpt *subopt = (hpair->subopt && hpair->subopt->path_count)?hpair->subopt:NULL;
    for(i = 0; i < span_n; i++){
        if(span_note[i].scored || (span_note[i].hpair != hpair))
            continue;
        span = span_note[i].span; /* one batch per Heuristic_Span (its two derived models) */
        n = 0;
        for(j = i; j < span_n; j++)
            if(!span_note[j].scored && (span_note[j].hpair == hpair) && (span_note[j].span == span)){
                src[n] = &span_note[j].src;
                dst[n] = &span_note[j].dst;
                which[n++] = j;
                }
        b200_span_scores(span, n, src, dst, hpair->user_data, hpair->subopt, scores);
        for(j = 0; j < n; j++){
            register B200_SpanNote *sn = &span_note[which[j]];
            sn->scored = TRUE;
            sn->score = scores[j];
            for(k = 0; k < 2; k++)
                sn->nb[k] = b200_blocked_list(subopt, k?&sn->dst:&sn->src, &sn->bq[k], &sn->bt[k]);
            }
        }
    g_free(src);
    g_free(dst);
    g_free(which);
    g_free(scores);
    return;
    }

void b200_BSDP_initialise(BSDP *bsdp, C4_Score threshold){
    register HPair *hpair = bsdp->user_data; /* BSDP_create(..., hpair), hpair.c:327-335 */
    register gint i, j, n;
    register Region **regions;
    register Optimal *optimal;
    if(noted_n){
        regions = g_new(Region*, noted_n);
        for(i = 0; i < noted_n; i++){
            if(!noted[i].optimal || (noted[i].hpair != hpair))
                continue;
            optimal = noted[i].optimal; /* one batch per derived model */
            n = 0;
            for(j = i; j < noted_n; j++)
                if((noted[j].hpair == hpair) && (noted[j].optimal == optimal)){
                    regions[n++] = noted[j].region;
                    noted[j].optimal = NULL;
                    }
            b200_prefetch_scores(optimal->find_score, n, regions, hpair->user_data, hpair->subopt);
            }
        g_free(regions);
        noted_n = 0; /* SARs of a comparison that never reached this point are dropped with it */
        }
    if(span_n){
        /* drop the notes of earlier comparisons (their SAR_Spans are gone), keep this graph's */
        register gint keep = 0;
        for(i = 0; i < span_n; i++){
            if(span_note[i].hpair == hpair){
                if(keep != i)
                    span_note[keep] = span_note[i];
                keep++;
            } else {
                g_free(span_note[i].bq[0]); g_free(span_note[i].bt[0]);
                g_free(span_note[i].bq[1]); g_free(span_note[i].bt[1]);
                }
            }
        span_n = keep;
        if(g_getenv("EXONERATE_B200_BSDP_SPANS") && !atoi(g_getenv("EXONERATE_B200_BSDP_SPANS")))
            span_notes_clear();
        else
            prefetch_spans(hpair);
        }
    BSDP_initialise(bsdp, threshold);
    return;
    }
