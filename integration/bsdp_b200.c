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
 * reported alignment has since touched simply misses and runs synchronously).  Span edges
 * (sar.c:898-917: two fills coupled through the integration matrices) still run per call.
 *
 * EXONERATE_B200_BSDP_BATCH=0 switches the prefetch off. */
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
    BSDP_initialise(bsdp, threshold);
    return;
    }
