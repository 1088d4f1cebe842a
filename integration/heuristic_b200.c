/* heuristic_b200.c -- binding of Heuristic_Span_integrate (src/bsdp/heuristic.c:589-678) to
 * c4b_span_integrate (include/c4b200.h).  Same prototype as src/bsdp/heuristic.h:163-164.
 *
 * COMPILED (integration/Makefile builds the object, so the binding in INTEGRATION.md section 6
 * is known to build against the reference's headers) but NOT LINKED into exonerate_b200: one
 * synchronous device call per span pair costs more in launch latency than the CPU scan it
 * replaces; it is meant to be fused into the batched SAR pass.  Linking
 * it would take the same objcopy --redefine-sym step on heuristic.o that hspset.o gets.
 *
 * Our own code; it includes the reference's headers because it implements its interface. */
#include "heuristic.h"
#include "c4b200.h"

c4b_engine *exonerate_b200_engine(void); /* viterbi_b200.c */

void B200_Heuristic_Span_integrate(Heuristic_Span *hs, Region *src, Region *dst){
    register gint i, j;
    register gint sq = src->query_length, st = src->target_length,
                  dq = dst->query_length, dt = dst->target_length;
    gint32 sreg[4], dreg[4], span[4];
    register gint32 *scores = g_new(gint32, (gsize)(sq+1)*(st+1)),
                    *pos = g_new(gint32, 2*(gsize)(dq+1)*(dt+1));
    sreg[0] = src->query_start; sreg[1] = src->target_start; sreg[2] = sq; sreg[3] = st;
    dreg[0] = dst->query_start; dreg[1] = dst->target_start; dreg[2] = dq; dreg[3] = dt;
    span[0] = hs->span->min_query;  span[1] = hs->span->max_query;
    span[2] = hs->span->min_target; span[3] = hs->span->max_target;
    for(i = 0; i <= sq; i++)
        for(j = 0; j <= st; j++)
            scores[(gsize)i*(st+1)+j] = hs->src_integration_matrix[i][j][0];
    if(c4b_span_integrate(exonerate_b200_engine(), scores, sreg, dreg, span, pos))
        g_error("libc4b200: %s", c4b_last_error());
    for(i = 0; i <= dq; i++)
        for(j = 0; j <= dt; j++){
            hs->dst_integration_matrix[i][j].query_pos = pos[2*((gsize)i*(dt+1)+j)];
            hs->dst_integration_matrix[i][j].target_pos = pos[2*((gsize)i*(dt+1)+j)+1];
            }
    g_free(scores);
    g_free(pos);
    return;
    }
