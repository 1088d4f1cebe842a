/* b200_binding.h -- what the objects of the reference-side binding share
 * (viterbi_b200.c, gam_b200.c, hspset_b200.c).  Our own code; see INTEGRATION.md. */
#ifndef B200_BINDING_H
#define B200_BINDING_H

#include "viterbi.h"
#include "match.h"
#include "subopt.h"
#include "c4b200.h"

/* the process-wide engine (EXONERATE_B200_DEVICE selects the CUDA ordinal) */
c4b_engine *exonerate_b200_engine(void);

/* EXONERATE_B200_DEVICES=all | 0,1,2,...: the exhaustive batches of the CLI are sharded over
 * these GPUs (c4b_group); NULL when unset (one device) */
c4b_group *exonerate_b200_group(void);

/* closed C4_Model of a Viterbi -> flat tables (built once per Viterbi) */
c4b_model *b200_tables_for(Viterbi *viterbi);
gboolean b200_model_has_splice(c4b_model *m);
void b200_fill_scoring(Match_ArgumentSet *mas, c4b_scoring *sc);
/* the four splice-site score arrays of a flattened target, int32[4][tlen]
 * (what intron_init_func prepares, src/model/intron.c:259-293); g_free() it */
gint32 *b200_splice_arrays(gchar *tseq, gint tlen);
/* SubOpt_Index (src/c4/subopt.c:250-338) of `region` as two arrays sorted by
 * (target_pos, query_pos), region coordinates; returns the count, 0 => both NULL */
gint b200_blocked_list(SubOpt *subopt, Region *region, gint32 **bq, gint32 **bt);

/* ---- the batch hook (gam_b200.c): answers computed ahead of the replay ------
 * While gam_b200.c replays a queued exhaustive comparison through the reference's own
 * GAM_Result_exhaustive_create, b200_replay points at that comparison's prefetched
 * FIND_PATH answers: round r = the r-th Optimal_find_path call of its --subopt loop
 * (src/hub/gam.c:1160-1172).  Viterbi_calculate takes an answer only if the call is the
 * one it was computed for (same Viterbi, sequences, full region, same blocked cells). */
typedef struct {
    gint n_blocked;
    gint32 *bq, *bt;
    c4b_result result;
    gint32 *ops; /* 2 * result.n_ops ints, (transition id, length) */
} B200_Round;

typedef struct {
    Viterbi *viterbi;
    Sequence *query, *target;
    B200_Round *rounds;
    gint n_rounds, cursor;
} B200_Replay;

/* ---- BSDP prefetch (bsdp_b200.c): FIND_SCORE answers of a comparison's SAR fills ----
 * b200_prefetch_scores() runs `n` region fills of ONE Viterbi (a terminal or join model of the
 * heuristic, src/bsdp/heuristic.c:242-325) on the current (query, target) as one device batch
 * and keeps the answers; Viterbi_calculate returns a kept answer only for exactly the same call
 * (same Viterbi, sequences, region and SubOpt blocked cells).  Dropped when the comparison changes. */
void b200_prefetch_scores(Viterbi *viterbi, gint n, Region **regions, gpointer user_data,
                          SubOpt *subopt);
/* n span edges of ONE Heuristic_Span (src / dst derived models of src/bsdp/heuristic.c:445-528) on
 * the current (query, target): what SAR_Span_find_score's dst Optimal_find_score would return,
 * scores[k], from one c4b_span_score_batch call.  bq/bt lists: the SubOpt blocked cells of each
 * region at this moment (kept by the caller to validate a later use), as b200_blocked_list returns. */
void b200_span_scores(gpointer heuristic_span, gint n, Region **src_regions, Region **dst_regions,
                      gpointer user_data, SubOpt *subopt, C4_Score *scores);
extern glong b200_stat_score_hits, b200_stat_score_prefetched, b200_stat_score_batches;

extern B200_Replay *b200_replay;
extern glong b200_stat_prefetch_hits, b200_stat_prefetch_misses;

#endif /* B200_BINDING_H */
