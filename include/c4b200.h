/* c4b200.h -- C ABI of the B200-native C4 Viterbi engine (libc4b200.so).
 *
 * Drop-in boundary for exonerate's C4 dynamic-programming hot path.  Plain C,
 * pointers + sizes only; no torch / CUDA types.  Every entry point names the
 * reference interface it replaces (paths relative to the reference tree).
 *
 *   reference                                   this ABI
 *   ------------------------------------------  -----------------------------
 *   closed C4_Model (src/c4/c4.h:172-194,       c4b_model (flat POD tables,
 *     C4_Model_close src/c4/c4.c:1669-1680)       transitions in closed order)
 *   C4_Calc callbacks (src/c4/c4.h:75-86,       c4b_calc {kind,param,protect}
 *     src/comparison/match.c:271-364,508-540,     evaluated on device
 *     src/model/affine.c:88-124,
 *     src/model/intron.c:138-161,
 *     src/model/phase.c:135-208)
 *   Submat / Translate (src/sequence/submat.h:  c4b_scoring
 *     31-56, src/sequence/translate.h:41-79)
 *   Optimal_find_score (src/c4/optimal.c:123)   c4b_find_score_batch
 *   Optimal_find_path  (src/c4/optimal.c:368)   c4b_find_path_batch
 *   Viterbi_DP_Func    (src/c4/viterbi.h:90-98) c4b_viterbi_calculate
 *   SubOpt_Index       (src/c4/subopt.h:55-80)  c4b_pair.blocked_*
 *   Alignment op list  (src/c4/alignment.h:     c4b_result.n_ops + ops[]
 *     34-50, Alignment_add alignment.c:75-100)    (transition id, run length)
 *
 * Errors: every call returns 0 on success, non-zero on failure; the message is
 * available from c4b_last_error().  The reference's g_error()->abort path
 * (src/general/argument.c:290-318) is the host shim's job.
 * There is no CPU fallback: without a CUDA device c4b_engine_create fails.
 */
#ifndef C4B200_H
#define C4B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C4B_ABI_VERSION 1

/* C4_Score is a 32-bit int; minus infinity as in src/c4/c4.h:28-30. */
typedef int32_t c4b_score;
#define C4B_IMPOSSIBLY_LOW_SCORE (-987654321)
#define C4B_IMPOSSIBLY_HIGH_SCORE (987654321)

#define C4B_MAX_STATES 32
#define C4B_MAX_TRANSITIONS 64
#define C4B_MAX_CALCS 32
#define C4B_MAX_SHADOW_SLOTS 4
#define C4B_SUBMAT_N 24

/* C4_Scope, src/c4/c4.h:91-97 (same numeric values). */
enum { C4B_SCOPE_ANYWHERE = 0, C4B_SCOPE_EDGE = 1, C4B_SCOPE_QUERY = 2,
       C4B_SCOPE_TARGET = 3, C4B_SCOPE_CORNER = 4 };

/* C4_Label, src/c4/c4.h:114-124 (same numeric values). */
enum { C4B_LABEL_NONE = 0, C4B_LABEL_MATCH = 1, C4B_LABEL_GAP = 2,
       C4B_LABEL_NER = 3, C4B_LABEL_5SS = 4, C4B_LABEL_3SS = 5,
       C4B_LABEL_INTRON = 6, C4B_LABEL_SPLIT_CODON = 7,
       C4B_LABEL_FRAMESHIFT = 8 };

/* C4_Protect, src/c4/c4.h:69-73. */
enum { C4B_PROTECT_NONE = 0, C4B_PROTECT_OVERFLOW = 1, C4B_PROTECT_UNDERFLOW = 2 };

/* Device forms of the reference's C4_CalcFunc callbacks.  "qp"/"tp" are the
 * SOURCE coordinates of the transition (src/c4/viterbi.c:749-752). */
enum {
    /* param[0]: gap open/extend, codon gap, frameshift penalties
     * (src/model/affine.c:88-124, src/model/frameshift.c:50-60) */
    C4B_CALC_CONST = 0,
    /* Submat_lookup(dna, q[qp], t[tp])  (src/comparison/match.c:271-285) */
    C4B_CALC_MATCH_DNA = 1,
    /* Submat_lookup(protein, q[qp], t[tp])  (match.c:287-295) */
    C4B_CALC_MATCH_PROTEIN = 2,
    /* protein q[qp] vs translated t[tp..tp+2]  (match.c:332-355) */
    C4B_CALC_MATCH_1_3 = 3,
    /* translated q[qp..qp+2] vs protein t[tp]  (match.c, 3:1 mirror) */
    C4B_CALC_MATCH_3_1 = 4,
    /* both codons translated  (match.c:508-540) */
    C4B_CALC_MATCH_3_3 = 5,
    /* param[0] + splice[param[1]][tp]: intron open at a splice site
     * (src/model/intron.c:138-161, is_pre) */
    C4B_CALC_SPLICE_PRE = 6,
    /* len = tp - slot + 2 outside [min_intron,max_intron] -> -inf, else
     * splice[param[1]][tp]; slot = shadow slot param[2] of the SOURCE cell
     * (intron.c:151-159; viterbi.c:747-748) */
    C4B_CALC_SPLICE_POST = 7,
    /* split codon, 1 base before the intron: slot>=1 ?
     * protein(q[qp]) vs codon(t[slot-1], t[tp], t[tp+1]) : -inf
     * (src/model/phase.c:135-208) */
    C4B_CALC_PHASE1_POST = 8,
    /* split codon, 2 bases before the intron: slot>=2 ?
     * protein(q[qp]) vs codon(t[slot-2], t[slot-1], t[tp]) : -inf */
    C4B_CALC_PHASE2_POST = 9,
    C4B_CALC_KIND_TOTAL
};

/* Index of the per-target splice score arrays (src/sequence/splice.h:47-52). */
enum { C4B_SPLICE_5_FORWARD = 0, C4B_SPLICE_3_FORWARD = 1,
       C4B_SPLICE_5_REVERSE = 2, C4B_SPLICE_3_REVERSE = 3, C4B_SPLICE_TOTAL = 4 };

typedef struct {
    int32_t kind;     /* C4B_CALC_* */
    int32_t protect;  /* C4B_PROTECT_* bits */
    int32_t param[4]; /* kind specific, see above */
} c4b_calc;

typedef struct {
    int32_t input;          /* source state id */
    int32_t output;         /* destination state id */
    int32_t advance_query;  /* >= 0 */
    int32_t advance_target; /* >= 0 */
    int32_t calc;           /* index into calcs, -1 = NULL calc (score 0) */
    int32_t label;          /* C4B_LABEL_* */
} c4b_transition;

/* A CLOSED model: ids assigned, transitions in the order produced by
 * C4_Model_topological_sort (src/c4/c4.c:1418-1486) -- that order IS the
 * tie-break contract (src/c4/viterbi.c:766-775) -- shadows designated
 * (c4.c:1638-1667). */
typedef struct {
    int32_t n_states;
    int32_t n_transitions;
    int32_t n_calcs;
    int32_t n_shadow_slots; /* total_shadow_designations */
    int32_t start_state, end_state;
    int32_t start_scope, end_scope; /* C4B_SCOPE_* */
    int32_t max_query_advance, max_target_advance;
    /* shadow start: when a transition whose INPUT is state s wins, slot d of
     * the source cell is stamped before transport (viterbi.c:413-422):
     * 0 = not a source of slot d, 1 = stamp source target position,
     * 2 = stamp source query position (sequence coordinates). */
    uint8_t shadow_start[C4B_MAX_STATES][C4B_MAX_SHADOW_SLOTS];
    c4b_transition transitions[C4B_MAX_TRANSITIONS];
    c4b_calc calcs[C4B_MAX_CALCS];
} c4b_model;

/* Flattened Submat + Translate + intron window (values the calcs read). */
typedef struct {
    int32_t dna_matrix[C4B_SUBMAT_N * C4B_SUBMAT_N];
    int32_t protein_matrix[C4B_SUBMAT_N * C4B_SUBMAT_N];
    uint8_t dna_index[256];     /* Submat.index; 24 = not in alphabet */
    uint8_t protein_index[256];
    uint8_t nt2d[256];          /* Translate.nt2d */
    uint8_t codon_aa[4096];     /* Translate.aa[Translate.trans[x]] */
    int32_t min_intron, max_intron; /* src/model/intron.c:24-32 */
} c4b_scoring;

/* One query x target lattice (an OPair / a Region of it). Sequences are raw
 * symbol bytes as returned by Sequence_strncpy (src/sequence/sequence.c:588).
 * The lattice is the Region {query_start,target_start,query_length,
 * target_length} (src/c4/region.h:26-32) of the two sequences. */
typedef struct {
    const uint8_t *query;
    const uint8_t *target;
    int32_t query_len, target_len;         /* whole sequence lengths */
    int32_t query_start, target_start;     /* region origin */
    int32_t query_length, target_length;   /* region extent */
    /* per-target splice-site scores, int32[target_len] each, or NULL when the
     * model has no SPLICE calcs (SplicePredictor_predict_array_int,
     * src/sequence/splice.c:383-397) */
    const int32_t *splice[C4B_SPLICE_TOTAL];
    /* SubOpt_Index (src/c4/subopt.c:250-338): blocked MATCH destination cells
     * in REGION coordinates, sorted by (target_pos, query_pos). May be NULL. */
    const int32_t *blocked_query_pos;
    const int32_t *blocked_target_pos;
    int32_t n_blocked;
    int32_t reserved; /* flags: C4B_PAIR_* below; 0 = none */
} c4b_pair;

/* c4b_pair.reserved bit: the query / target / splice buffers of this pair stay valid and
 * unchanged until c4b_engine_forget_buffers(): the engine may keep its device copies and find
 * them again by (address, length).  For the reference's heuristic mode, which asks for
 * thousands of small region fills on one (query, target) (src/bsdp/sar.c): each fill then
 * uploads its own descriptors only.  Blocked-cell lists are per call and never kept. */
#define C4B_PAIR_BUFFERS_STABLE 1
/* c4b_pair.reserved bit: the query and target buffers are page-locked host memory (cudaHostAlloc /
 * cudaHostRegister, e.g. a pinned framework tensor).  When EVERY pair of an affine-family batch
 * says so, the engine DMAs straight from them instead of copying through its own pinned bounce
 * buffer (equally long sequences at a constant stride -- rows of one array -- as one 2-D copy).
 * The buffers must stay valid until the batch has been fetched. */
#define C4B_PAIR_BUFFERS_PINNED 2

/* Result of one lattice.  Coordinates are SEQUENCE coordinates like
 * Alignment.region (src/c4/alignment.h:39-45); ops index into ops[] buffers as
 * (transition id, run length) pairs in path order, RLE-merged exactly as
 * Alignment_add does (src/c4/alignment.c:75-100). */
typedef struct {
    c4b_score score;
    int32_t query_start, target_start; /* alignment region start */
    int32_t query_end, target_end;     /* alignment region end (exclusive) */
    int32_t n_ops;                     /* number of (transition,length) pairs */
    int64_t ops_offset;                /* first pair = ops[2*ops_offset] */
    int32_t status;                    /* 0 ok, 1 below threshold (no path) */
    int32_t reserved;
} c4b_result;

typedef struct c4b_engine c4b_engine;
typedef struct c4b_batch c4b_batch;

/* The device compiler of c4b_model_specialise includes this header for the types
 * above only. */
#ifndef C4B200_TYPES_ONLY

/* ---- engine ------------------------------------------------------------ */
int c4b_abi_version(void);
const char *c4b_last_error(void);
/* device = CUDA ordinal. Fails (no fallback) when no CUDA device is usable.
 * Threading: the reference's host side is single-threaded (no USE_PTHREADS; module-level
 * statics in match.c / argument.c), and so is an engine: use it, and the batches made from
 * it, from one host thread at a time; engines on different devices (one process per GPU, or
 * one thread per engine) are independent.  c4b_last_error() is per thread. */
int c4b_engine_create(int device, c4b_engine **out);
void c4b_engine_destroy(c4b_engine *e);
/* Use an existing CUDA stream (cudaStream_t as void*) instead of the engine's
 * own; lets a host framework order our launches with its copies. */
int c4b_engine_set_stream(c4b_engine *e, void *cuda_stream);
/* Drop every device copy kept under C4B_PAIR_BUFFERS_STABLE (call before freeing or
 * rewriting such a buffer; batches created from such pairs must be destroyed first).
 * Waits for the engine's stream. */
void c4b_engine_forget_buffers(c4b_engine *e);
/* The same for the copies of ONE host buffer (by address), e.g. when a host keeps a target and
 * its splice arrays across comparisons and only the query changes. */
void c4b_engine_forget_buffer(c4b_engine *e, const void *host);
/* Counters since engine creation: kernels launched by this library. */
int64_t c4b_engine_kernel_launches(const c4b_engine *e);

/* ---- batched Optimal_* ------------------------------------------------- */
/* Optimal_find_score over n independent lattices. Host buffers in, host out. */
int c4b_find_score_batch(c4b_engine *e, const c4b_model *model,
                         const c4b_scoring *scoring, int32_t n,
                         const c4b_pair *pairs, c4b_score *scores);

/* Optimal_find_path over n independent lattices: score, alignment region and
 * the operation list.  ops must hold 2*ops_capacity int32; if the paths need
 * more, the call fails with an error naming the required capacity.
 * threshold as in Optimal_find_path (status=1 when score < threshold). */
int c4b_find_path_batch(c4b_engine *e, const c4b_model *model,
                        const c4b_scoring *scoring, int32_t n,
                        const c4b_pair *pairs, c4b_score threshold,
                        c4b_result *results, int32_t *ops, int64_t ops_capacity);

/* ---- resident batches (upload once, run many, fetch) ------------------- */
/* Stages sequences / tables in HBM.  want_path: 0 = scores only. */
int c4b_batch_create(c4b_engine *e, const c4b_model *model,
                     const c4b_scoring *scoring, int32_t n,
                     const c4b_pair *pairs, int want_path, c4b_batch **out);
/* Enqueue the fill (+ traceback) kernels on the engine stream; asynchronous. */
int c4b_batch_run(c4b_batch *b, c4b_score threshold);
/* Wait, then copy results (and ops when want_path) to the host. */
int c4b_batch_fetch(c4b_batch *b, c4b_result *results, int32_t *ops,
                    int64_t ops_capacity);
/* Number of (transition,length) pairs the paths of the last run hold in total = the
 * ops_capacity c4b_batch_fetch needs (an Alignment's operation_list lengths summed,
 * src/c4/alignment.h:34-50).  Waits for the run; -1 on error.  Lets a host size the ops
 * buffer exactly instead of for the worst case (query_length + target_length per lattice). */
int64_t c4b_batch_ops_needed(c4b_batch *b);
/* Device pointer to the c4b_result[n] array of the last run (pair order), valid
 * until the batch is destroyed; for device-side consumers such as an NCCL
 * gather of the per-pair records.  NULL for score-only batches before a run. */
const void *c4b_batch_device_results(const c4b_batch *b);
/* Lattice cells (sum of query_length*target_length) of the batch. */
int64_t c4b_batch_cells(const c4b_batch *b);
/* Device time of the dominant fill kernel of the last run, ms (CUDA events on
 * the engine stream); <0 if not run. */
double c4b_batch_last_fill_ms(c4b_batch *b);
/* Name of the kernel path chosen for this batch ("affine_systolic", "generic"). */
const char *c4b_batch_kernel_name(const c4b_batch *b);
/* One line on how the batch was routed: how many lattices took the packed 16-bit, int32 and
 * table-driven kernels, rows per lane, warps per lattice (diagnostics; bench.py prints it). */
const char *c4b_batch_description(const c4b_batch *b);
void c4b_batch_destroy(c4b_batch *b);

/* ---- device groups: one batch over several GPUs of one box ----------------
 * Pairs are independent (GAM_Result_exhaustive_create touches per-pair state only,
 * src/hub/gam.c:1140-1180), so a batch shards with no exchange step inside the DP: the group owns
 * one engine and one host thread per device, deals the lattices to the devices by cost
 * (query_length x target_length, largest first, to the least loaded device), runs the shards
 * concurrently -- each device is sent only its own shard's sequences -- and merges results and
 * op lists back into pair order.  Results are identical to c4b_find_path_batch on one device.
 * A single lattice is never split across devices.  devices = CUDA ordinals; n_devices = 0 means
 * every visible device.  Same single-host-thread rule as an engine. */
typedef struct c4b_group c4b_group;
int c4b_group_create(int n_devices, const int *devices, c4b_group **out);
void c4b_group_destroy(c4b_group *g);
int c4b_group_size(const c4b_group *g);
/* Optimal_find_score / Optimal_find_path over n lattices, sharded.  The path variant returns the
 * op list in a buffer of exactly the needed size, allocated by the library (release it with
 * c4b_free); results[k].ops_offset index into it. */
int c4b_group_find_score_batch(c4b_group *g, const c4b_model *model, const c4b_scoring *scoring,
                               int32_t n, const c4b_pair *pairs, c4b_score *scores);
int c4b_group_find_path_batch(c4b_group *g, const c4b_model *model, const c4b_scoring *scoring,
                              int32_t n, const c4b_pair *pairs, c4b_score threshold,
                              c4b_result *results, int32_t **ops_out, int64_t *n_ops_out);
void c4b_free(void *p);
/* kernels launched by all engines of the group since creation */
int64_t c4b_group_kernel_launches(const c4b_group *g);

/* ---- single-lattice Viterbi_DP_Func shape ------------------------------ */
/* mode: 0 FIND_SCORE, 1 FIND_PATH, 2 FIND_REGION (src/c4/viterbi.h:104-109).
 * One synchronous lattice; what Bootstrapper_lookup()'s trampolines call. */
int c4b_viterbi_calculate(c4b_engine *e, const c4b_model *model,
                          const c4b_scoring *scoring, const c4b_pair *pair,
                          int mode, c4b_result *result, int32_t *ops,
                          int64_t ops_capacity);

/* ---- BSDP derived models with cell callbacks -------
 * The heuristic path scores derived models whose START / END states carry callbacks
 * (C4_Model_configure_start_state / _end_state, src/c4/c4.h; used by
 * src/bsdp/heuristic.c only):
 *   cell_end_func(cell, cell_size, query_pos, target_pos, user_data)   viterbi.c:792-797
 *     Heuristic_Bound_report_end_func (heuristic.c:139-145, bound fills :150-207) and
 *     Heuristic_Span_src_report_end_func (:385-410) -- both only READ END's cell of every
 *     lattice cell that reaches END;
 *   cell_start_func(query_pos, target_pos, user_data) -> cell              viterbi.c:727-741
 *     Heuristic_Span_dst_init_start_func (:412-443) -- START's score and shadow slots per cell.
 * Callbacks cannot run on the device, and they do not need to: the binding evaluates
 * cell_start_func for every cell of the region beforehand (start_cells) and calls
 * cell_end_func afterwards on what the device returns (end_cells).  Both tables are
 * (query_length+1) x (target_length+1) cells (query-major) of 1 + n_shadow_slots ints;
 * end_cells entries of cells that never reach END are left untouched; either may be NULL.
 * mode 0 FIND_SCORE, 1 FIND_PATH; one synchronous lattice, table-driven kernel. */
int c4b_viterbi_calculate_cells(c4b_engine *e, const c4b_model *model, const c4b_scoring *scoring,
                                const c4b_pair *pair, int mode, const c4b_score *start_cells,
                                c4b_score *end_cells, c4b_result *result, int32_t *ops,
                                int64_t ops_capacity);

/* ---- HSP seeding / extension ----------------------
 * Replaces the per-seed work of HSPset_seed_hsp (src/comparison/hspset.c:933-997):
 * HSP_trim_ends (:850-878), HSP_init (:725-745), HSP_extend with masking forbidden
 * and then ignored (:747-812), the threshold test of HSP_store (:893-894) and
 * HSP_find_cobs (:426-441), for EVERY seed of one query x target comparison in one
 * launch.  The diagonal horizon (:951-972,991-996) is sequential state over the seed
 * list; it only decides which seeds are looked at, so the caller (our hspset.c
 * binding, exonerate_b200.engine.HSPset) replays it over these results in seed order.
 * match_kind: C4B_CALC_MATCH_DNA (advance 1,1), C4B_CALC_MATCH_PROTEIN (1,1),
 * C4B_CALC_MATCH_1_3 (protein query vs translated DNA target, advance 1,3),
 * C4B_CALC_MATCH_3_1 (translated DNA query vs protein target, 3,1) or
 * C4B_CALC_MATCH_3_3 (both translated: codon2codon, 3,3) -- Match_Type_* of match.h:45-51.
 * query_mask / target_mask: one byte per position, non-zero = masked
 * (Alphabet_is_masked, src/sequence/alphabet.h:87-88); NULL = nothing masked. */
typedef struct {
    int32_t match_kind;
    int32_t seedlen;   /* HSP_Param.seedlen = wordlen / query advance (hspset.c:110-117) */
    int32_t dropoff;   /* HSP_Param.dropoff */
    int32_t threshold; /* HSP_Param.threshold */
} c4b_hsp_param;

typedef struct {
    int32_t query_start, target_start;
} c4b_hsp_seed;

typedef struct {
    int32_t query_start, target_start;
    int32_t length;     /* match-state visits */
    int32_t score;
    int32_t cobs;       /* centre offset by score; valid when stored */
    int32_t stored;     /* 1: score >= threshold (HSP_store keeps it); 0: dropped */
    int32_t target_end; /* what the seed writes into horizon[0] (hspset.c:985,995) */
    int32_t status;     /* 0 ok; 1 = "Initial HSP score less than zero" (g_error in the reference) */
} c4b_hsp;

int c4b_hsp_extend_batch(c4b_engine *e, const c4b_scoring *scoring, const c4b_hsp_param *param,
                         const uint8_t *query, int32_t query_len, const uint8_t *query_mask,
                         const uint8_t *target, int32_t target_len, const uint8_t *target_mask,
                         int32_t n_seeds, const c4b_hsp_seed *seeds, c4b_hsp *out);

/* ---- BSDP span integration -----------------------------
 * Replaces the scan of Heuristic_Span_integrate (src/bsdp/heuristic.c:589-678): for every cell
 * of the dst region, the position of the best src-region START-side score a span of
 * [min_query,max_query] x [min_target,max_target] symbols can bridge, first in (query, target)
 * order on ties (:638), or (-1,-1).  src_scores = src_integration_matrix[x][y][0] as
 * (src_ql+1) x (src_tl+1) ints; regions are {query_start, target_start, query_length,
 * target_length} (region.h:26-32); span = {min_query, max_query, min_target, max_target}
 * (C4_Span, c4.h:160-170); positions = (dst_ql+1) x (dst_tl+1) x {query_pos, target_pos}
 * (Heuristic_Span_Cell, heuristic.h:75-78).  One synchronous call. */
int c4b_span_integrate(c4b_engine *e, const c4b_score *src_scores, const int32_t *src_region,
                       const int32_t *dst_region, const int32_t *span, int32_t *positions);

/* ---- BSDP span edges, batched ------------------------------------------
 * SAR_Span_find_score (src/bsdp/sar.c:898-917) for n span edges of one heuristic comparison in a
 * few launches: the src fills (src_model: START at the region's corner, END's cell reported from
 * every cell, heuristic.c:385-410), Heuristic_Span_integrate (:589-678) and the START table of
 * Heuristic_Span_dst_init_start_func (:412-443) built on the device from the src END cells, the
 * dst fills (dst_model: START's cell from that table, END at the corner).  scores[k] = what the
 * dst Optimal_find_score returns (the caller subtracts the SAR components).  span = {min_query,
 * max_query, min_target, max_target} (C4_Span, c4.h:160-170). */
typedef struct {
    c4b_pair src, dst;
    int32_t span[4];
} c4b_span_job;
int c4b_span_score_batch(c4b_engine *e, const c4b_model *src_model, const c4b_model *dst_model,
                         const c4b_scoring *scoring, int32_t n, const c4b_span_job *jobs, c4b_score *scores);

/* ---- model specialisation ---------------------------------------------------
 * Device counterpart of the reference's per-model code generation (Viterbi_compile /
 * Codegen, src/c4/viterbi.c:1638-1727, src/c4/codegen.c; archived by the bootstrapper,
 * src/c4/bootstrapper.c): the table-driven fill is compiled for ONE closed model at run
 * time (NVRTC, sm_100a) and cached.  The batch entry points do this themselves for large
 * batches (env C4B_GENERIC_JIT: 0 never, 1 always); this call only compiles -- every variant
 * the launcher can ask for -- and, when C4B_JIT_CACHE_DIR is set, leaves the cubins there: a
 * host can fill the cache (the reference's "bootstrapper" step) or check that a model
 * specialises without a GPU.
 * mode: 0 FIND_SCORE, 1 FIND_PATH, 2 FIND_REGION; cta_threads: 128, 256 or 512.
 * Returns 0 and the cubin size, or -1 with c4b_last_error(). */
int c4b_model_specialise(const c4b_model *model, int32_t mode, int32_t cta_threads, int64_t *cubin_bytes);

#endif /* C4B200_TYPES_ONLY */

#ifdef __cplusplus
}
#endif
#endif /* C4B200_H */
