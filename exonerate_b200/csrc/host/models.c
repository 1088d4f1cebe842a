/* models.c -- the shipped alignment models, built on c4model.c.
 * Follows src/model/{ungapped,affine,frameshift,intron,phase,est2genome,
 * protein2dna,protein2genome,coding2coding,modeltype}.c: same states,
 * transitions, labels, advances and INSERTION ORDER (the closed transition
 * order -- the tie-break contract -- is a function of it).
 */
#include "c4host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void C4_Params_default(C4_Params *p) {
    p->gap_open = -12;
    p->gap_extend = -4;
    p->codon_gap_open = -18;
    p->codon_gap_extend = -8;
    p->frameshift = -28;
    p->intron_open = -30;
    p->min_intron = 30;
    p->max_intron = 200000;
    p->match_max_dna = 5;       /* nucleic */
    p->match_max_protein = 11;  /* blosum62 */
}

static char *join(const char *a, const char *b, const char *c) {
    size_t n = strlen(a) + strlen(b) + (c ? strlen(c) : 0) + 1;
    char *r = (char *)malloc(n);
    strcpy(r, a);
    strcat(r, b);
    if (c) strcat(r, c);
    return r;
}

/* src/comparison/match.c:72-125 */
Match_Type Match_Type_find(Alphabet_Type q, Alphabet_Type t, int translate_both) {
    if (q == Alphabet_Type_DNA && t == Alphabet_Type_DNA)
        return translate_both ? Match_Type_CODON2CODON : Match_Type_DNA2DNA;
    if (q == Alphabet_Type_PROTEIN && t == Alphabet_Type_PROTEIN) return Match_Type_PROTEIN2PROTEIN;
    if (q == Alphabet_Type_DNA) return Match_Type_DNA2PROTEIN;
    return Match_Type_PROTEIN2DNA;
}

static const struct {
    const char *name;
    int aq, at, kind;
} match_info[] = {
    {"dna2dna", 1, 1, C4B_CALC_MATCH_DNA},      {"protein2protein", 1, 1, C4B_CALC_MATCH_PROTEIN},
    {"dna2protein", 3, 1, C4B_CALC_MATCH_3_1},  {"protein2dna", 1, 3, C4B_CALC_MATCH_1_3},
    {"codon", 3, 3, C4B_CALC_MATCH_3_3},
};

/* src/model/ungapped.c:106-166 */
C4_Model *Ungapped_create(Match_Type mt, const C4_Params *p) {
    char *name = join("ungapped:", match_info[mt].name, NULL);
    C4_Model *m = C4_Model_create(name);
    C4_State *match = C4_Model_add_state(m, "match");
    C4_Calc *calc = C4_Model_add_calc(m, "match",
                                      mt == Match_Type_DNA2DNA ? p->match_max_dna : p->match_max_protein,
                                      match_info[mt].kind, NULL, C4_Protect_NONE);
    free(name);
    C4_Model_add_transition(m, "start to match", NULL, match, 0, 0, NULL, C4_Label_NONE);
    C4_Model_add_transition(m, "match to end", match, NULL, 0, 0, NULL, C4_Label_NONE);
    C4_Model_add_transition(m, "match", match, match, match_info[mt].aq, match_info[mt].at, calc, C4_Label_MATCH);
    C4_Model_close(m);
    return m;
}

static const char *affine_type_name(Affine_Model_Type type) {
    switch (type) {
    case Affine_Model_Type_GLOBAL: return "global";
    case Affine_Model_Type_BESTFIT: return "bestfit";
    case Affine_Model_Type_LOCAL: return "local";
    default: return "overlap";
    }
}

/* src/model/affine.c:150-255 */
C4_Model *Affine_create(Affine_Model_Type type, Alphabet_Type qt, Alphabet_Type tt, int translate_both,
                        const C4_Params *p) {
    const Match_Type mt = Match_Type_find(qt, tt, translate_both);
    C4_Model *m = Ungapped_create(mt, p);
    C4_Scope scope = C4_Scope_ANYWHERE;
    C4_State *ins, *del, *match_in, *match_out;
    C4_Transition *match;
    C4_Calc *open, *extend;
    int32_t par[4] = {0, 0, 0, 0};
    int aq, at, codon;
    char *name = join("affine:", affine_type_name(type), ":");
    char *full = join(name, match_info[mt].name, NULL);
    switch (type) {
    case Affine_Model_Type_GLOBAL: scope = C4_Scope_CORNER; break;
    case Affine_Model_Type_BESTFIT: scope = C4_Scope_QUERY; break;
    case Affine_Model_Type_LOCAL: scope = C4_Scope_ANYWHERE; break;
    case Affine_Model_Type_OVERLAP: scope = C4_Scope_EDGE; break;
    }
    C4_Model_rename(m, full);
    free(name);
    free(full);
    C4_Model_configure_start_state(m, scope);
    C4_Model_configure_end_state(m, scope);
    C4_Model_open(m);
    ins = C4_Model_add_state(m, "insert");
    del = C4_Model_add_state(m, "delete");
    match = C4_Model_select_single_transition(m, C4_Label_MATCH);
    match_in = C4_Transition_input(match);
    match_out = C4_Transition_output(match);
    aq = match_info[mt].aq;
    at = match_info[mt].at;
    /* codon penalties when the match advances by 3 (affine.c:198-209); the reported
     * max_score stays the plain penalty (affine.c:210-217) */
    codon = (aq > at ? aq : at) == 3;
    par[0] = codon ? p->codon_gap_open : p->gap_open;
    open = C4_Model_add_calc(m, "gap open", p->gap_open, C4B_CALC_CONST, par, C4_Protect_NONE);
    par[0] = codon ? p->codon_gap_extend : p->gap_extend;
    extend = C4_Model_add_calc(m, "gap extend", p->gap_extend, C4B_CALC_CONST, par, C4_Protect_NONE);
    C4_Model_add_transition(m, "match to insert", match_in, ins, aq, 0, open, C4_Label_GAP);
    C4_Model_add_transition(m, "match to delete", match_in, del, 0, at, open, C4_Label_GAP);
    C4_Model_add_transition(m, "insert", ins, ins, aq, 0, extend, C4_Label_GAP);
    C4_Model_add_transition(m, "insert to match", ins, match_out, 0, 0, NULL, C4_Label_NONE);
    C4_Model_add_transition(m, "delete", del, del, 0, at, extend, C4_Label_GAP);
    C4_Model_add_transition(m, "delete to match", del, match_out, 0, 0, NULL, C4_Label_NONE);
    /* "match portal" on the match calc (affine.c:243-246) */
    C4_Model_add_portal(m, "match portal", C4_Transition_calc(match), aq, at);
    C4_Model_close(m);
    return m;
}

/* src/model/frameshift.c:82-126 */
void Frameshift_add(C4_Model *m, C4_State *match_state, const char *suffix, int on_query, const C4_Params *p) {
    char *sname = join("frameshift ", suffix, NULL), *name;
    C4_State *fs = C4_Model_add_state(m, sname);
    /* reuse the model's frameshift calc if it has one (frameshift.c:64-75) */
    C4_Calc *calc = C4_Model_find_calc(m, "frameshift");
    int32_t par[4] = {0, 0, 0, 0};
    if (!calc) {
        par[0] = p->frameshift;
        calc = C4_Model_add_calc(m, "frameshift", p->frameshift, C4B_CALC_CONST, par, C4_Protect_NONE);
    }
    name = join("frameshift open 1 ", suffix, NULL);
    C4_Model_add_transition(m, name, match_state, fs, on_query ? 1 : 0, on_query ? 0 : 1, calc, C4_Label_FRAMESHIFT);
    free(name);
    name = join("frameshift open 2 ", suffix, NULL);
    C4_Model_add_transition(m, name, match_state, fs, on_query ? 2 : 0, on_query ? 0 : 2, calc, C4_Label_FRAMESHIFT);
    free(name);
    name = join("frameshift close 0 ", suffix, NULL);
    C4_Model_add_transition(m, name, fs, match_state, 0, 0, NULL, C4_Label_NONE);
    free(name);
    name = join("frameshift close 3 ", suffix, NULL);
    C4_Model_add_transition(m, name, fs, match_state, on_query ? 3 : 0, on_query ? 0 : 3, NULL, C4_Label_FRAMESHIFT);
    free(name);
    free(sname);
}

/* src/model/intron.c:496-586,588-697.  Only target introns have a device form
 * (the shipped est2genome / protein2genome use nothing else). */
C4_Model *Intron_create(const char *suffix, int on_query, int on_target, int is_forward, const C4_Params *p) {
    char *name = join("intron ", suffix, NULL), *tmp;
    C4_Model *m = C4_Model_create(name);
    C4_State *intron;
    C4_Calc *pre, *post;
    const char *pre_name = is_forward ? "5'ss forward" : "3'ss reverse";
    const char *post_name = is_forward ? "3'ss forward" : "5'ss reverse";
    const C4_Label pre_label = is_forward ? C4_Label_5SS : C4_Label_3SS;
    const C4_Label post_label = is_forward ? C4_Label_3SS : C4_Label_5SS;
    const int qa = on_query ? 2 : 0, ta = on_target ? 2 : 0;
    int32_t par[4] = {0, 0, 0, 0};
    if (on_query) {
        fprintf(stderr, "c4host: query introns have no device calc form\n");
        abort();
    }
    /* Intron_add_calc: pre = open penalty + splice score, post = length window + splice
     * score, both UNDERFLOW-protected (intron.c:571-579) */
    tmp = join(pre_name, " ", suffix);
    par[0] = p->intron_open;
    par[1] = is_forward ? C4B_SPLICE_5_FORWARD : C4B_SPLICE_3_REVERSE;
    pre = C4_Model_add_calc(m, tmp, 0, C4B_CALC_SPLICE_PRE, par, C4_Protect_UNDERFLOW);
    free(tmp);
    tmp = join(post_name, " ", suffix);
    par[0] = 0;
    par[1] = is_forward ? C4B_SPLICE_3_FORWARD : C4B_SPLICE_5_REVERSE;
    post = C4_Model_add_calc(m, tmp, 0, C4B_CALC_SPLICE_POST, par, C4_Protect_UNDERFLOW);
    free(tmp);
    intron = C4_Model_add_state(m, name);
    tmp = join("(START) to ", name, NULL);
    C4_Model_add_transition(m, tmp, NULL, intron, qa, ta, pre, pre_label);
    free(tmp);
    if (on_target) {
        tmp = join("target intron loop ", suffix, NULL);
        C4_Model_add_transition(m, tmp, intron, intron, 0, 1, NULL, C4_Label_INTRON);
        free(tmp);
    }
    tmp = join(name, " to (END)", NULL);
    C4_Model_add_transition(m, tmp, intron, NULL, qa, ta, post, post_label);
    free(tmp);
    tmp = join("intron span", suffix, NULL);
    C4_Model_add_span(m, tmp, intron, 0, 0, p->min_intron, p->max_intron);
    free(tmp);
    tmp = join("target intron ", suffix, NULL);
    C4_Model_add_shadow(m, tmp, NULL, NULL, C4_ShadowStart_TARGET_POS);
    free(tmp);
    free(name);
    C4_Model_close(m);
    return m;
}

/* src/model/phase.c:354-547 (protein vs genomic target: on_query = 0, on_target = 1) */
C4_Model *Phase_create(const char *suffix, Match_Type mt, int on_query, int on_target, const C4_Params *p) {
    char full[128], buf[192], iname[160];
    C4_Model *m, *i00, *i12, *i21;
    C4_Calc *c1, *c2;
    C4_State *p1pre, *p1post, *p2pre, *p2post;
    C4_Transition *t1post, *t2post;
    if (mt != Match_Type_PROTEIN2DNA || on_query || !on_target) {
        fprintf(stderr, "c4host: only the protein2dna target phase model has device calcs\n");
        abort();
    }
    snprintf(full, sizeof(full), "phase%s%s%s%s%s", suffix ? " " : "", suffix ? suffix : "", suffix ? " " : "",
             on_query ? "Q" : "-", on_target ? "T" : "-");
    m = C4_Model_create(full);
    snprintf(iname, sizeof(iname), "0:0 %s", full);
    i00 = Intron_create(iname, on_query, on_target, 1, p);
    iname[0] = '1'; iname[2] = '2';
    i12 = Intron_create(iname, on_query, on_target, 1, p);
    iname[0] = '2'; iname[2] = '1';
    i21 = Intron_create(iname, on_query, on_target, 1, p);
    snprintf(buf, sizeof(buf), "phase1post to dst %s", full);
    c1 = C4_Model_add_calc(m, buf, p->match_max_protein, C4B_CALC_PHASE1_POST, NULL, C4_Protect_NONE);
    snprintf(buf, sizeof(buf), "phase2post to dst %s", full);
    c2 = C4_Model_add_calc(m, buf, p->match_max_protein, C4B_CALC_PHASE2_POST, NULL, C4_Protect_NONE);
    snprintf(buf, sizeof(buf), "phase1pre %s", full);
    p1pre = C4_Model_add_state(m, buf);
    snprintf(buf, sizeof(buf), "phase1post %s", full);
    p1post = C4_Model_add_state(m, buf);
    snprintf(buf, sizeof(buf), "phase2pre %s", full);
    p2pre = C4_Model_add_state(m, buf);
    snprintf(buf, sizeof(buf), "phase2post %s", full);
    p2post = C4_Model_add_state(m, buf);
    /* against a peptide, on the target: pre1 (0,1) post1 (1,2) pre2 (0,2) post2 (1,1) */
    snprintf(buf, sizeof(buf), "(START) to phase1pre %s", full);
    C4_Model_add_transition(m, buf, NULL, p1pre, 0, 1, NULL, C4_Label_SPLIT_CODON);
    snprintf(buf, sizeof(buf), "(START) to phase2pre %s", full);
    C4_Model_add_transition(m, buf, NULL, p2pre, 0, 2, NULL, C4_Label_SPLIT_CODON);
    snprintf(buf, sizeof(buf), "phase1post %s to (END)", full);
    t1post = C4_Model_add_transition(m, buf, p1post, NULL, 1, 2, c1, C4_Label_SPLIT_CODON);
    snprintf(buf, sizeof(buf), "phase2post %s to (END)", full);
    t2post = C4_Model_add_transition(m, buf, p2post, NULL, 1, 1, c2, C4_Label_SPLIT_CODON);
    C4_Model_insert(m, i00, NULL, NULL);
    C4_Model_insert(m, i12, p1pre, p1post);
    C4_Model_insert(m, i21, p2pre, p2post);
    {   /* the intron shadows of the split-codon paths also end on the post
         * transitions, whose calcs read the intron start (phase.c:525-536) */
        C4_Shadow_add_dst_transition(C4_Model_shadow_at(m, 1), t1post);
        C4_Shadow_add_dst_transition(C4_Model_shadow_at(m, 2), t2post);
    }
    C4_Model_destroy(i00);
    C4_Model_destroy(i12);
    C4_Model_destroy(i21);
    C4_Model_close(m);
    return m;
}

/* src/model/est2genome.c:57-93 */
C4_Model *EST2Genome_create(const C4_Params *p) {
    C4_Model *m = Affine_create(Affine_Model_Type_LOCAL, Alphabet_Type_DNA, Alphabet_Type_DNA, 0, p);
    C4_Model *fwd, *rev;
    C4_Transition *match[2];
    C4_Model_rename(m, "est2genome");
    C4_Model_open(m);
    C4_Model_make_stereo(m, "forward", "reverse");
    C4_Model_select_transitions(m, C4_Label_MATCH, match, 2);
    fwd = Intron_create("forward", 0, 1, 1, p);
    rev = Intron_create("reverse", 0, 1, 0, p);
    C4_Model_insert(m, fwd, C4_Transition_input(match[0]), C4_Transition_input(match[0]));
    C4_Model_insert(m, rev, C4_Transition_input(match[1]), C4_Transition_input(match[1]));
    C4_Model_destroy(fwd);
    C4_Model_destroy(rev);
    C4_Model_close(m);
    return m;
}

/* src/model/protein2dna.c:55-72 */
C4_Model *Protein2DNA_create(Affine_Model_Type type, const C4_Params *p) {
    C4_Model *m = Affine_create(type, Alphabet_Type_PROTEIN, Alphabet_Type_DNA, 0, p);
    char *name = join("protein2dna:", affine_type_name(type), NULL);
    C4_Transition *match;
    C4_Model_rename(m, name);
    free(name);
    C4_Model_open(m);
    match = C4_Model_select_single_transition(m, C4_Label_MATCH);
    Frameshift_add(m, C4_Transition_input(match), "p2d", 0, p);
    C4_Model_close(m);
    return m;
}

/* src/model/protein2genome.c:45-68 */
C4_Model *Protein2Genome_create(Affine_Model_Type type, const C4_Params *p) {
    C4_Model *m = Protein2DNA_create(type, p), *phase;
    char *name = join("protein2genome:", affine_type_name(type), NULL);
    C4_Transition *match;
    C4_Model_rename(m, name);
    free(name);
    C4_Model_open(m);
    match = C4_Model_select_single_transition(m, C4_Label_MATCH);
    phase = Phase_create(NULL, Match_Type_PROTEIN2DNA, 0, 1, p);
    C4_Model_insert(m, phase, C4_Transition_input(match), C4_Transition_output(match));
    C4_Model_destroy(phase);
    C4_Model_close(m);
    return m;
}

/* src/model/coding2coding.c:50-66 */
C4_Model *Coding2Coding_create(const C4_Params *p) {
    C4_Model *m = Affine_create(Affine_Model_Type_LOCAL, Alphabet_Type_DNA, Alphabet_Type_DNA, 1, p);
    C4_Transition *match;
    C4_Model_rename(m, "coding2coding");
    C4_Model_open(m);
    match = C4_Model_select_single_transition(m, C4_Label_MATCH);
    Frameshift_add(m, C4_Transition_input(match), "query", 1, p);
    Frameshift_add(m, C4_Transition_input(match), "target", 0, p);
    C4_Model_close(m);
    return m;
}

/* src/model/modeltype.c:83-123,225-290 */
C4_Model *Model_Type_get_model(const char *name, Alphabet_Type qt, Alphabet_Type tt, const C4_Params *p) {
    static const char *names[][2] = {
        {"ungapped", "u"}, {"affine:global", "a:g"}, {"affine:bestfit", "a:b"}, {"affine:local", "a:l"},
        {"affine:overlap", "a:o"}, {"est2genome", "e2g"}, {"protein2dna", "p2d"},
        {"protein2dna:bestfit", "p2d:b"}, {"protein2genome", "p2g"}, {"protein2genome:bestfit", "p2g:b"},
        {"coding2coding", "c2c"}};
    int k, which = -1;
    for (k = 0; k < 11; k++)
        if (!strcmp(name, names[k][0]) || !strcmp(name, names[k][1])) which = k;
    switch (which) {
    case 0: return Ungapped_create(Match_Type_find(qt, tt, 0), p);
    case 1: return Affine_create(Affine_Model_Type_GLOBAL, qt, tt, 0, p);
    case 2: return Affine_create(Affine_Model_Type_BESTFIT, qt, tt, 0, p);
    case 3: return Affine_create(Affine_Model_Type_LOCAL, qt, tt, 0, p);
    case 4: return Affine_create(Affine_Model_Type_OVERLAP, qt, tt, 0, p);
    case 5: return EST2Genome_create(p);
    case 6: return Protein2DNA_create(Affine_Model_Type_LOCAL, p);
    case 7: return Protein2DNA_create(Affine_Model_Type_BESTFIT, p);
    case 8: return Protein2Genome_create(Affine_Model_Type_LOCAL, p);
    case 9: return Protein2Genome_create(Affine_Model_Type_BESTFIT, p);
    case 10: return Coding2Coding_create(p);
    default: return NULL; /* Model_Type_from_string g_error()s; the caller reports */
    }
}

/* ---- one-call entry for bindings: model name -> engine tables ---------------- */
int c4b_host_model(const char *name, int query_is_protein, int target_is_protein, const C4_Params *params,
                   c4b_model *out, char **description) {
    C4_Params def;
    C4_Model *m;
    int rc;
    if (!params) {
        C4_Params_default(&def);
        params = &def;
    }
    m = Model_Type_get_model(name, query_is_protein ? Alphabet_Type_PROTEIN : Alphabet_Type_DNA,
                             target_is_protein ? Alphabet_Type_PROTEIN : Alphabet_Type_DNA, params);
    if (!m) return -2;
    rc = C4_Model_flatten(m, out);
    if (description) *description = C4_Model_describe(m);
    C4_Model_destroy(m);
    return rc;
}
