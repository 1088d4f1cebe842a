/* gam_b200.c -- the batch hook behind the exonerate CLI.
 *
 * The reference compares one (query, target) pair at a time: Analysis_Pair_compare ->
 * Analysis_ExhaustiveJob_run -> GAM_Result_exhaustive_create -> OPair_next_path ->
 * Optimal_find_path -> Viterbi_calculate (src/hub/analysis.c:196-249, src/hub/gam.c:1140-1180,
 * src/c4/opair.c:42-56).  One synchronous lattice per call cannot fill a GPU (SURVEY.md 8b
 * "needed addition").  This object batches the pairs WITHOUT touching src/hub:
 *
 *   analysis.o is linked as a copy in which two external references are renamed by objcopy
 *   (integration/Makefile; the trick hspset_b200.c already uses):
 *       GAM_Result_exhaustive_create -> b200_GAM_Result_exhaustive_create
 *       GAM_report                   -> b200_GAM_report        (analysis.c:1421)
 *   gam.o stays the reference's own, unrenamed.
 *
 *   b200_GAM_Result_exhaustive_create only ENQUEUES {gam, query, target} (ref-counted, like
 *   Analysis_ExhaustiveJob_create) and returns NULL, so the caller submits nothing.
 *   A FLUSH -- queue full, or b200_GAM_report at the end of Analysis_process -- then
 *     1. flattens every distinct sequence once (Sequence_strncpy; sequences are virtual),
 *     2. runs Optimal_find_path's one FIND_PATH fill of ALL queued pairs as one
 *        c4b_find_path batch (round 0), and, with --subopt yes, the following iterations of
 *        the sub-optimal series as further batches (round r blocks the match cells of the
 *        paths of rounds < r, built with the reference's own SubOpt_add_alignment),
 *     3. replays the pairs IN ARRIVAL ORDER through the reference's real
 *        GAM_Result_exhaustive_create / GAM_Result_submit / GAM_Result_destroy: its
 *        Viterbi_calculate calls find their answers prefetched (viterbi_b200.c:
 *        prefetch_lookup) and return without a device call.
 *   Thresholds (--score, --percent, --bestn), the sub-optimal loop, refinement, result
 *   storage and printing are the reference's own code running in its own order, so stdout
 *   stays byte-identical; a call the prefetch did not anticipate (e.g. after --refine changed
 *   the alignment) simply runs synchronously as before.
 *
 * Our own code; it includes the reference's headers because it implements the reference's
 * interface.  Knobs: EXONERATE_B200_BATCH=0 disables batching (one lattice per call, as in
 * round 1); EXONERATE_B200_BATCH_PAIRS / _BATCH_MB / _BATCH_GCELLS bound the queue;
 * EXONERATE_B200_SUBOPT_ROUNDS bounds the prefetched sub-optimal iterations. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "gam.h"
#include "optimal.h"
#include "alignment.h"
#include "intron.h"
#include "b200_binding.h"

typedef struct { /* a flattened sequence, shared by every queued pair that uses it */
    guint64 hash;
    gint len;
    gchar *flat;
    gint32 *splice; /* 4 x int32[len] for targets of models with splice calcs */
} B200_Flat;

typedef struct {
    GAM *gam;
    Sequence *query, *target;
    B200_Flat *qflat, *tflat;
    SubOpt *subopt;      /* our own copy of the state GAM_Result's subopt will go through */
    B200_Round *rounds;
    gint n_rounds;
    gboolean active;     /* still producing alignments at or above the static threshold */
} B200_Job;

static struct {
    B200_Job *job;
    gint n, cap;
    gint64 cells;
    gsize bytes;
    B200_Flat **flat; /* open-addressing table by content hash */
    gint flat_cap, flat_n;
} queue = {NULL, 0, 0, 0, 0, NULL, 0, 0};

static glong stat_flushes = 0, stat_pairs = 0, stat_rounds = 0, stat_lattices = 0;
static gdouble stat_flatten = 0, stat_device = 0, stat_replay = 0, stat_splice = 0;
static gint64 stat_cells = 0;

static gdouble now_seconds(void){
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9*ts.tv_nsec;
    }

static void print_stats(void){
    fprintf(stderr, "exonerate_b200: batch hook: %ld pair(s) in %ld flush(es), %ld device "
                    "round(s), %ld lattices, %.4g Gcells; flatten %.3f s, splice arrays %.3f s, "
                    "device %.3f s (%.1f GCUPS), replay %.3f s\n",
            stat_pairs, stat_flushes, stat_rounds, stat_lattices, stat_cells*1e-9,
            stat_flatten, stat_splice, stat_device,
            (stat_device > 0)?stat_cells*1e-9/stat_device:0.0, stat_replay);
    }

static glong env_long(const gchar *name, glong dflt){
    register const gchar *v = g_getenv(name);
    return v?atol(v):dflt;
    }

/* ---- flattened sequences, one copy per distinct content ------------------------- */
static guint64 hash_bytes(const gchar *s, gint len){
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

static void flat_table_grow(void){
    register gint i, old_cap = queue.flat_cap, slot;
    register B200_Flat **old = queue.flat;
    queue.flat_cap = old_cap?old_cap*2:1024;
    queue.flat = g_new0(B200_Flat*, queue.flat_cap);
    for(i = 0; i < old_cap; i++)
        if(old[i]){
            slot = old[i]->hash & (queue.flat_cap-1);
            while(queue.flat[slot])
                slot = (slot+1) & (queue.flat_cap-1);
            queue.flat[slot] = old[i];
            }
    g_free(old);
    return;
    }

static B200_Flat *flat_fetch(Sequence *s){
    register gchar *buf = g_new(gchar, s->len+16);
    register guint64 h;
    register gint slot;
    register B200_Flat *f;
    Sequence_strncpy(s, 0, s->len, buf);
    memset(buf+s->len, 0, 16);
    h = hash_bytes(buf, s->len);
    if(2*(queue.flat_n+1) > queue.flat_cap)
        flat_table_grow();
    slot = h & (queue.flat_cap-1);
    while((f = queue.flat[slot])){
        if((f->hash == h) && (f->len == s->len) && !memcmp(f->flat, buf, s->len)){
            g_free(buf); /* same content seen before (all-vs-all re-reads every target) */
            return f;
            }
        slot = (slot+1) & (queue.flat_cap-1);
        }
    f = g_new0(B200_Flat, 1);
    f->hash = h;
    f->len = s->len;
    f->flat = buf;
    queue.flat[slot] = f;
    queue.flat_n++;
    queue.bytes += s->len;
    return f;
    }

static void flat_table_clear(void){
    register gint i;
    for(i = 0; i < queue.flat_cap; i++)
        if(queue.flat[i]){
            g_free(queue.flat[i]->flat);
            g_free(queue.flat[i]->splice);
            g_free(queue.flat[i]);
            queue.flat[i] = NULL;
            }
    queue.flat_n = 0;
    queue.bytes = 0;
    return;
    }

/* ---- one device round over the active jobs ---------------------------------------- */
static void run_round(GAM *gam, gint round, gint *active, gint n_active){
    register Viterbi *viterbi = gam->optimal->find_path;
    register c4b_model *tables = b200_tables_for(viterbi);
    register c4b_group *group = exonerate_b200_group(); /* several GPUs: EXONERATE_B200_DEVICES */
    register c4b_engine *engine = group?NULL:exonerate_b200_engine();
    register c4b_pair *pairs = g_new0(c4b_pair, n_active);
    register c4b_result *results = g_new0(c4b_result, n_active);
    gint32 *ops = NULL;
    register gint k, i;
    gint64 need;
    register gboolean with_splice = b200_model_has_splice(tables);
    register B200_Job *job;
    register B200_Round *rd;
    register gdouble t0;
    c4b_batch *batch = NULL;
    c4b_scoring scoring;
    b200_fill_scoring(Match_ArgumentSet_create(NULL), &scoring);
    for(k = 0; k < n_active; k++){
        job = &queue.job[active[k]];
        job->rounds = g_renew(B200_Round, job->rounds, round+1);
        rd = &job->rounds[round];
        memset(rd, 0, sizeof(B200_Round));
        if(round > 0){
            Region full;
            full.ref_count = -1; /* static region (region.h:27) */
            full.query_start = full.target_start = 0;
            full.query_length = job->query->len;
            full.target_length = job->target->len;
            rd->n_blocked = b200_blocked_list(job->subopt, &full, &rd->bq, &rd->bt);
            }
        pairs[k].query = (const uint8_t*)job->qflat->flat;
        pairs[k].target = (const uint8_t*)job->tflat->flat;
        pairs[k].query_len = pairs[k].query_length = job->query->len;
        pairs[k].target_len = pairs[k].target_length = job->target->len;
        pairs[k].blocked_query_pos = rd->bq;
        pairs[k].blocked_target_pos = rd->bt;
        pairs[k].n_blocked = rd->n_blocked;
        if(with_splice)
            for(i = 0; i < 4; i++)
                pairs[k].splice[i] = job->tflat->splice + (gsize)i*job->target->len;
        stat_cells += (gint64)job->query->len * job->target->len;
        }
    t0 = now_seconds();
    /* create / run / size the op buffer exactly / fetch: the worst-case op count
     * (query + target per lattice) of a 10k-pair batch would be gigabytes */
    if(group){ /* lattices dealt to the devices by cost, shards run concurrently, merged in pair order */
        if(c4b_group_find_path_batch(group, tables, &scoring, n_active, pairs,
                                     C4B_IMPOSSIBLY_LOW_SCORE, results, &ops, &need))
            g_error("libc4b200: %s", c4b_last_error());
    } else {
        if(c4b_batch_create(engine, tables, &scoring, n_active, pairs, 1, &batch)
        || c4b_batch_run(batch, C4B_IMPOSSIBLY_LOW_SCORE))
            g_error("libc4b200: %s", c4b_last_error());
        need = c4b_batch_ops_needed(batch);
        if(need < 0)
            g_error("libc4b200: %s", c4b_last_error());
        ops = g_new(gint32, 2*need+2);
        if(c4b_batch_fetch(batch, results, ops, need))
            g_error("libc4b200: %s", c4b_last_error());
        c4b_batch_destroy(batch);
        }
    stat_device += now_seconds() - t0;
    for(k = 0; k < n_active; k++){
        job = &queue.job[active[k]];
        rd = &job->rounds[round];
        rd->result = results[k];
        rd->ops = g_new(gint32, 2*results[k].n_ops+2);
        memcpy(rd->ops, ops + 2*results[k].ops_offset, 2*results[k].n_ops*sizeof(gint32));
        rd->result.ops_offset = 0;
        job->n_rounds = round+1;
        }
    stat_rounds++;
    stat_lattices += n_active;
    if(group)
        c4b_free(ops);
    else
        g_free(ops);
    g_free(results);
    g_free(pairs);
    return;
    }

/* the path of the last round joins the job's SubOpt exactly as GAM_Result_add_alignment
 * will do it during the replay (gam.c:659-676: SubOpt_add_alignment of the Alignment that
 * Viterbi_Data_create_Alignment builds from these ops) */
static void job_block_last_path(B200_Job *job, C4_Model *model){
    register B200_Round *rd = &job->rounds[job->n_rounds-1];
    register Region *region = Region_create(rd->result.query_start, rd->result.target_start,
        rd->result.query_end - rd->result.query_start,
        rd->result.target_end - rd->result.target_start);
    register Alignment *alignment = Alignment_create(model, region, rd->result.score);
    register gint i;
    for(i = 0; i < rd->result.n_ops; i++)
        Alignment_add(alignment, model->transition_list->pdata[rd->ops[2*i]], rd->ops[2*i+1]);
    if(!job->subopt)
        job->subopt = SubOpt_create(job->query->len, job->target->len);
    SubOpt_add_alignment(job->subopt, alignment);
    Alignment_destroy(alignment);
    Region_destroy(region);
    return;
    }

static void flush(void){
    register gint k, r, n_active, max_rounds;
    register gint *active;
    register B200_Job *job;
    register GAM *gam;
    register GAM_Result *gam_result;
    register gdouble t0;
    register gboolean with_splice;
    B200_Replay replay;
    if(!queue.n)
        return;
    gam = queue.job[0].gam; /* every queued job has the same GAM (enqueue flushes on change) */
    with_splice = b200_model_has_splice(b200_tables_for(gam->optimal->find_path));
    t0 = now_seconds();
    for(k = 0; k < queue.n; k++){
        job = &queue.job[k];
        job->qflat = flat_fetch(job->query);
        job->tflat = flat_fetch(job->target);
        }
    stat_flatten += now_seconds() - t0;
    if(with_splice){
        t0 = now_seconds();
        for(k = 0; k < queue.n; k++){
            job = &queue.job[k];
            if(!job->tflat->splice)
                job->tflat->splice = b200_splice_arrays(job->tflat->flat, job->tflat->len);
            }
        stat_splice += now_seconds() - t0;
        }
    /* rounds of the sub-optimal series.  Round r+1 of a pair is worth computing only if its
     * round-r path can be accepted at all: the reference's loop stops at the first score
     * below the threshold, which is never lower than --score (gam.c:679-731). */
    max_rounds = gam->gas->use_subopt?(gint)env_long("EXONERATE_B200_SUBOPT_ROUNDS", 16):1;
    if(gam->gas->refinement != GAM_Refinement_NONE)
        max_rounds = 1; /* refinement replaces the path that gets blocked: cannot anticipate */
    active = g_new(gint, queue.n);
    for(k = 0; k < queue.n; k++){
        active[k] = k;
        queue.job[k].active = TRUE;
        }
    n_active = queue.n;
    for(r = 0; (r < max_rounds) && n_active; r++){
        run_round(gam, r, active, n_active);
        n_active = 0;
        if(r+1 >= max_rounds)
            break;
        for(k = 0; k < queue.n; k++){
            job = &queue.job[k];
            if(!job->active)
                continue;
            if((job->rounds[r].result.score < gam->gas->threshold)
            || !job->rounds[r].result.n_ops){
                job->active = FALSE;
                continue;
                }
            job_block_last_path(job, gam->optimal->find_path->model);
            active[n_active++] = k;
            }
        }
    g_free(active);
    /* replay, in arrival order, through the reference's own code */
    t0 = now_seconds();
    for(k = 0; k < queue.n; k++){
        job = &queue.job[k];
        replay.viterbi = gam->optimal->find_path;
        replay.query = job->query;
        replay.target = job->target;
        replay.rounds = job->rounds;
        replay.n_rounds = job->n_rounds;
        replay.cursor = 0;
        b200_replay = &replay;
        gam_result = GAM_Result_exhaustive_create(job->gam, job->query, job->target);
        b200_replay = NULL;
        if(gam_result){
            GAM_Result_submit(gam_result);
            GAM_Result_destroy(gam_result);
            }
        for(r = 0; r < job->n_rounds; r++){
            g_free(job->rounds[r].bq);
            g_free(job->rounds[r].bt);
            g_free(job->rounds[r].ops);
            }
        g_free(job->rounds);
        if(job->subopt)
            SubOpt_destroy(job->subopt);
        GAM_destroy(job->gam);
        Sequence_destroy(job->query);
        Sequence_destroy(job->target);
        }
    stat_replay += now_seconds() - t0;
    stat_flushes++;
    stat_pairs += queue.n;
    queue.n = 0;
    queue.cells = 0;
    flat_table_clear();
    return;
    }

/* ---- the two renamed entry points ---------------------------------------------------- */
GAM_Result *b200_GAM_Result_exhaustive_create(GAM *gam, Sequence *query, Sequence *target){
    static gint enabled = -1;
    static glong max_pairs, max_bytes;
    static gint64 max_cells;
    register B200_Job *job;
    if(enabled < 0){
        enabled = env_long("EXONERATE_B200_BATCH", 1)?1:0;
        /* defaults: a flush of 4096 pairs keeps the GPU busy for ~0.1 s at the metric shape while
         * the queued Sequences (the reference re-reads every target per query) stay cache- and
         * page-friendly: 10k pairs of 1 kbp x 100 kbp ran in 4.85 s with flushes of 2000 pairs
         * and 5.62 s as one flush (profiles/r02_cli.md) */
        max_pairs = env_long("EXONERATE_B200_BATCH_PAIRS", 4096);
        max_bytes = env_long("EXONERATE_B200_BATCH_MB", 1024) << 20;
        max_cells = (gint64)env_long("EXONERATE_B200_BATCH_GCELLS", 4000) * 1000000000ll;
        if(g_getenv("EXONERATE_B200_STATS"))
            atexit(print_stats);
        }
    if(!enabled)
        return GAM_Result_exhaustive_create(gam, query, target);
    if(queue.n && (queue.job[0].gam != gam))
        flush();
    if(queue.n == queue.cap){
        queue.cap = queue.cap?queue.cap*2:256;
        queue.job = g_renew(B200_Job, queue.job, queue.cap);
        }
    job = &queue.job[queue.n++];
    memset(job, 0, sizeof(B200_Job));
    job->gam = GAM_share(gam);
    job->query = Sequence_share(query);
    job->target = Sequence_share(target);
    queue.cells += (gint64)query->len * target->len;
    queue.bytes += query->len + target->len; /* upper bound until the flush dedupes */
    if((queue.n >= max_pairs) || (queue.cells >= max_cells) || (queue.bytes >= (gsize)max_bytes))
        flush();
    return NULL; /* nothing to submit yet: the flush submits in arrival order */
    }

void b200_GAM_report(GAM *gam){
    flush();
    GAM_report(gam);
    return;
    }
