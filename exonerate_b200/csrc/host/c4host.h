/* c4host.h -- host-side C4 model layer of the B200 engine (plain C99, no glib).
 *
 * Mirrors the part of the reference's C4 model API that produces the hot
 * path's "program": building an alignment automaton, CLOSING it (ids,
 * transition precedence order, shadow slot designation, max advances;
 * src/c4/c4.c:1349-1383,1418-1486,1513-1535,1638-1680) and the shipped model
 * builders of src/model/{ungapped,affine,intron,phase,frameshift,est2genome,
 * protein2dna,protein2genome,coding2coding}.c.  Same names and argument
 * meaning as src/c4/c4.h:198-303, with ONE deliberate difference: a C4_Calc
 * carries a device calc kind (include/c4b200.h C4B_CALC_*) instead of a host
 * function pointer -- there is no host callback on the CUDA path.
 *
 * The closed model is flattened to the c4b_model the engine consumes
 * (C4_Model_flatten).  Written from the reference's documented behaviour; the
 * closed forms are checked table-by-table against dumps of the reference's own
 * models (tests/test_host_models.py).
 */
#ifndef C4HOST_H
#define C4HOST_H

#include "c4b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct C4_State C4_State;
typedef struct C4_Calc C4_Calc;
typedef struct C4_Transition C4_Transition;
typedef struct C4_Shadow C4_Shadow;
typedef struct C4_Model C4_Model;

typedef enum { C4_Scope_ANYWHERE, C4_Scope_EDGE, C4_Scope_QUERY, C4_Scope_TARGET, C4_Scope_CORNER } C4_Scope;
typedef enum { C4_Label_NONE, C4_Label_MATCH, C4_Label_GAP, C4_Label_NER, C4_Label_5SS, C4_Label_3SS,
               C4_Label_INTRON, C4_Label_SPLIT_CODON, C4_Label_FRAMESHIFT } C4_Label;
typedef enum { C4_Protect_NONE = 0, C4_Protect_OVERFLOW = 1, C4_Protect_UNDERFLOW = 2 } C4_Protect;
typedef enum { C4_ShadowStart_TARGET_POS = 1, C4_ShadowStart_QUERY_POS = 2 } C4_ShadowStart;

/* ---- src/c4/c4.h:198-303 ------------------------------------------------ */
C4_Model *C4_Model_create(const char *name);           /* created open, with START and END */
void C4_Model_destroy(C4_Model *model);
void C4_Model_rename(C4_Model *model, const char *name);
void C4_Model_open(C4_Model *model);
void C4_Model_close(C4_Model *model);
C4_State *C4_Model_add_state(C4_Model *model, const char *name);
/* kind/param/protect: device form of the reference's calc_func (c4b200.h) */
C4_Calc *C4_Model_add_calc(C4_Model *model, const char *name, c4b_score max_score, int kind,
                           const int32_t param[4], C4_Protect protect);
/* NULL input = START, NULL output = END */
C4_Transition *C4_Model_add_transition(C4_Model *model, const char *name, C4_State *input,
                                       C4_State *output, int advance_query, int advance_target,
                                       C4_Calc *calc, C4_Label label);
/* NULL src = START; NULL dst = every transition into END */
C4_Shadow *C4_Model_add_shadow(C4_Model *model, const char *name, C4_State *src, C4_Transition *dst,
                               C4_ShadowStart start);
void C4_Shadow_add_src_state(C4_Shadow *shadow, C4_State *src);
void C4_Shadow_add_dst_transition(C4_Shadow *shadow, C4_Transition *dst);
void C4_Model_add_portal(C4_Model *model, const char *name, C4_Calc *calc, int advance_query,
                         int advance_target);
void C4_Model_add_span(C4_Model *model, const char *name, C4_State *span_state, int min_query,
                       int max_query, int min_target, int max_target);
void C4_Model_configure_start_state(C4_Model *model, C4_Scope scope);
void C4_Model_configure_end_state(C4_Model *model, C4_Scope scope);
C4_Transition *C4_Model_select_single_transition(C4_Model *model, C4_Label label);
int C4_Model_select_transitions(C4_Model *model, C4_Label label, C4_Transition **out, int max);
void C4_Model_make_stereo(C4_Model *model, const char *suffix_a, const char *suffix_b);
/* insert a CLOSED model between src and dst of an OPEN one (NULL = START / END) */
void C4_Model_insert(C4_Model *target, C4_Model *insert, C4_State *src, C4_State *dst);
C4_State *C4_Transition_input(C4_Transition *t);
C4_State *C4_Transition_output(C4_Transition *t);
C4_Calc *C4_Transition_calc(C4_Transition *t);
C4_Calc *C4_Model_find_calc(C4_Model *model, const char *name);   /* first calc of that name */
C4_Shadow *C4_Model_shadow_at(C4_Model *model, int index);        /* model->shadow_list->pdata[index] */

/* closed model -> engine tables; returns 0, or <0 with C4_host_error() set */
int C4_Model_flatten(const C4_Model *model, c4b_model *out);
/* text dump, one record per line, same fields as oracle/ref_driver.c's dump of
 * the reference model (without the macro strings); caller frees with free() */
char *C4_Model_describe(const C4_Model *model);
const char *C4_host_error(void);

/* ---- penalties the builders read (the reference's ArgumentSets) ---------- */
typedef struct {
    int32_t gap_open, gap_extend;             /* src/model/affine.c:24-29   (-12, -4) */
    int32_t codon_gap_open, codon_gap_extend; /* affine.c:30-35             (-18, -8) */
    int32_t frameshift;                       /* src/model/frameshift.c:23  (-28) */
    int32_t intron_open;                      /* src/model/intron.c:24-32   (-30) */
    int32_t min_intron, max_intron;           /*                            (30, 200000) */
    int32_t match_max_dna, match_max_protein; /* Match_max_score, only reported */
} C4_Params;
void C4_Params_default(C4_Params *p);

typedef enum { Alphabet_Type_DNA, Alphabet_Type_PROTEIN } Alphabet_Type;
typedef enum { Affine_Model_Type_GLOBAL, Affine_Model_Type_BESTFIT, Affine_Model_Type_LOCAL,
               Affine_Model_Type_OVERLAP } Affine_Model_Type;
typedef enum { Match_Type_DNA2DNA, Match_Type_PROTEIN2PROTEIN, Match_Type_DNA2PROTEIN,
               Match_Type_PROTEIN2DNA, Match_Type_CODON2CODON } Match_Type;

/* ---- src/model builders (closed models) ---------------------------------- */
Match_Type Match_Type_find(Alphabet_Type query_type, Alphabet_Type target_type, int translate_both);
C4_Model *Ungapped_create(Match_Type match_type, const C4_Params *p);
C4_Model *Affine_create(Affine_Model_Type type, Alphabet_Type query_type, Alphabet_Type target_type,
                        int translate_both, const C4_Params *p);
C4_Model *Intron_create(const char *suffix, int on_query, int on_target, int is_forward,
                        const C4_Params *p);
C4_Model *Phase_create(const char *suffix, Match_Type match_type, int on_query, int on_target,
                       const C4_Params *p);
void Frameshift_add(C4_Model *model, C4_State *match_state, const char *suffix, int apply_to_query,
                    const C4_Params *p);
C4_Model *EST2Genome_create(const C4_Params *p);
C4_Model *Protein2DNA_create(Affine_Model_Type type, const C4_Params *p);
C4_Model *Protein2Genome_create(Affine_Model_Type type, const C4_Params *p);
C4_Model *Coding2Coding_create(const C4_Params *p);
/* Model_Type_from_string + Model_Type_get_model (src/model/modeltype.c:48-98,225-290):
 * "ungapped", "affine:global|bestfit|local|overlap", "est2genome", "protein2dna",
 * "protein2dna:bestfit", "protein2genome", "protein2genome:bestfit", "coding2coding"
 * and the reference's short forms (u, a:g, a:b, a:l, a:o, e2g, p2d, p2g, c2c). */
C4_Model *Model_Type_get_model(const char *name, Alphabet_Type query_type, Alphabet_Type target_type,
                               const C4_Params *p);
/* one call for bindings: model name -> engine tables (+ optional text dump to free()).
 * params NULL = reference defaults. Returns 0, -2 unknown model, <0 C4_host_error(). */
int c4b_host_model(const char *name, int query_is_protein, int target_is_protein, const C4_Params *params,
                   c4b_model *out, char **description);

/* ---- splice-site score arrays (src/sequence/splice.c) ---------------------- */
/* SplicePredictor_predict_array_int over the whole sequence; type = C4B_SPLICE_* */
void c4b_host_splice_array(int type, const uint8_t *seq, int32_t len, int force_gtag, int32_t *out);
/* all four site types, C4B_SPLICE_* order, out[4*len] (feeds c4b_pair.splice[]) */
void c4b_host_splice_arrays(const uint8_t *seq, int32_t len, int force_gtag, int32_t *out);

#ifdef __cplusplus
}
#endif
#endif
