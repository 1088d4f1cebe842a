/* c4model.c -- C4 model graph: build, insert, stereo, CLOSE, flatten.
 * See c4host.h.  Each function names the reference behaviour it reproduces
 * (paths relative to the reference tree); the code itself is our own.
 */
#include "c4host.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- tiny growable pointer vector ---------------------------------------- */
typedef struct {
    void **v;
    int n, cap;
} Vec;
static void vec_push(Vec *a, void *p) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 8;
        a->v = (void **)realloc(a->v, sizeof(void *) * (size_t)a->cap);
    }
    a->v[a->n++] = p;
}
static void vec_free(Vec *a) {
    free(a->v);
    a->v = NULL;
    a->n = a->cap = 0;
}
static char *dup_str(const char *s) {
    size_t n = strlen(s) + 1;
    char *r = (char *)malloc(n);
    memcpy(r, s, n);
    return r;
}
static char *cat_str(const char *a, const char *b, const char *c) {
    size_t n = strlen(a) + strlen(b) + (c ? strlen(c) : 0) + 1;
    char *r = (char *)malloc(n);
    strcpy(r, a);
    strcat(r, b);
    if (c) strcat(r, c);
    return r;
}

static char host_error[256];
const char *C4_host_error(void) { return host_error; }
static void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(host_error, sizeof(host_error), fmt, ap);
    va_end(ap);
}

/* ---- graph objects (src/c4/c4.h:61-194) ----------------------------------- */
struct C4_State {
    char *name;
    int id;
    Vec in_tr, out_tr, src_shadows;
};
struct C4_Calc {
    char *name;
    int id;
    c4b_score max_score;
    c4b_calc dev;
};
struct C4_Transition {
    char *name;
    int id;
    C4_State *input, *output;
    int advance_query, advance_target;
    C4_Calc *calc;
    C4_Label label;
    Vec dst_shadows;
};
struct C4_Shadow {
    char *name;
    int id, designation;
    C4_ShadowStart start;
    Vec src_states, dst_transitions;
};
typedef struct {
    char *name;
    int id, advance_query, advance_target;
    C4_Calc *calc;
} Portal;
typedef struct {
    char *name;
    int id, min_query, max_query, min_target, max_target;
    C4_State *state;
} Span;
struct C4_Model {
    char *name;
    int is_open;
    Vec states, transitions, shadows, calcs, portals, spans;
    C4_State *start, *end;
    C4_Scope start_scope, end_scope;
    int max_query_advance, max_target_advance, total_shadow_designations;
};

C4_State *C4_Transition_input(C4_Transition *t) { return t->input; }
C4_State *C4_Transition_output(C4_Transition *t) { return t->output; }
C4_Calc *C4_Transition_calc(C4_Transition *t) { return t->calc; }
C4_Shadow *C4_Model_shadow_at(C4_Model *m, int index) {
    return (index >= 0 && index < m->shadows.n) ? (C4_Shadow *)m->shadows.v[index] : NULL;
}
C4_Calc *C4_Model_find_calc(C4_Model *m, const char *name) {
    int i;
    for (i = 0; i < m->calcs.n; i++)
        if (!strcmp(((C4_Calc *)m->calcs.v[i])->name, name)) return (C4_Calc *)m->calcs.v[i];
    return NULL;
}

/* ---- construction (c4.c:344-545) ------------------------------------------ */
C4_State *C4_Model_add_state(C4_Model *m, const char *name) {
    C4_State *s = (C4_State *)calloc(1, sizeof(*s));
    s->name = dup_str(name);
    vec_push(&m->states, s);
    return s;
}

C4_Model *C4_Model_create(const char *name) {
    C4_Model *m = (C4_Model *)calloc(1, sizeof(*m));
    m->name = dup_str(name);
    m->is_open = 1;
    m->start = C4_Model_add_state(m, "START");
    m->end = C4_Model_add_state(m, "END");
    m->start_scope = m->end_scope = C4_Scope_ANYWHERE;
    return m;
}

void C4_Model_rename(C4_Model *m, const char *name) {
    free(m->name);
    m->name = dup_str(name);
}

void C4_Model_destroy(C4_Model *m) {
    int i;
    for (i = 0; i < m->states.n; i++) {
        C4_State *s = (C4_State *)m->states.v[i];
        vec_free(&s->in_tr); vec_free(&s->out_tr); vec_free(&s->src_shadows);
        free(s->name); free(s);
    }
    for (i = 0; i < m->transitions.n; i++) {
        C4_Transition *t = (C4_Transition *)m->transitions.v[i];
        vec_free(&t->dst_shadows);
        free(t->name); free(t);
    }
    for (i = 0; i < m->shadows.n; i++) {
        C4_Shadow *s = (C4_Shadow *)m->shadows.v[i];
        vec_free(&s->src_states); vec_free(&s->dst_transitions);
        free(s->name); free(s);
    }
    for (i = 0; i < m->calcs.n; i++) {
        C4_Calc *c = (C4_Calc *)m->calcs.v[i];
        free(c->name); free(c);
    }
    for (i = 0; i < m->portals.n; i++) { free(((Portal *)m->portals.v[i])->name); free(m->portals.v[i]); }
    for (i = 0; i < m->spans.n; i++) { free(((Span *)m->spans.v[i])->name); free(m->spans.v[i]); }
    vec_free(&m->states); vec_free(&m->transitions); vec_free(&m->shadows);
    vec_free(&m->calcs); vec_free(&m->portals); vec_free(&m->spans);
    free(m->name);
    free(m);
}

C4_Calc *C4_Model_add_calc(C4_Model *m, const char *name, c4b_score max_score, int kind,
                           const int32_t param[4], C4_Protect protect) {
    C4_Calc *c = (C4_Calc *)calloc(1, sizeof(*c));
    int k;
    c->name = dup_str(name);
    c->max_score = max_score;
    c->dev.kind = kind;
    c->dev.protect = (int32_t)protect;
    for (k = 0; k < 4; k++) c->dev.param[k] = param ? param[k] : 0;
    vec_push(&m->calcs, c);
    return c;
}

C4_Transition *C4_Model_add_transition(C4_Model *m, const char *name, C4_State *input, C4_State *output,
                                       int aq, int at, C4_Calc *calc, C4_Label label) {
    C4_Transition *t = (C4_Transition *)calloc(1, sizeof(*t));
    t->name = dup_str(name);
    t->input = input ? input : m->start;
    t->output = output ? output : m->end;
    t->advance_query = aq;
    t->advance_target = at;
    t->calc = calc;
    t->label = label;
    vec_push(&t->input->out_tr, t);
    vec_push(&t->output->in_tr, t);
    vec_push(&m->transitions, t);
    return t;
}

void C4_Shadow_add_src_state(C4_Shadow *sh, C4_State *src) {
    vec_push(&sh->src_states, src);
    vec_push(&src->src_shadows, sh);
}
void C4_Shadow_add_dst_transition(C4_Shadow *sh, C4_Transition *dst) {
    vec_push(&sh->dst_transitions, dst);
    vec_push(&dst->dst_shadows, sh);
}

/* c4.c:450-483: NULL dst means every transition into END */
C4_Shadow *C4_Model_add_shadow(C4_Model *m, const char *name, C4_State *src, C4_Transition *dst,
                               C4_ShadowStart start) {
    C4_Shadow *sh = (C4_Shadow *)calloc(1, sizeof(*sh));
    int i;
    sh->name = dup_str(name);
    sh->start = start;
    sh->designation = -1;
    C4_Shadow_add_src_state(sh, src ? src : m->start);
    if (dst) {
        C4_Shadow_add_dst_transition(sh, dst);
    } else {
        for (i = 0; i < m->end->in_tr.n; i++)
            C4_Shadow_add_dst_transition(sh, (C4_Transition *)m->end->in_tr.v[i]);
    }
    vec_push(&m->shadows, sh);
    return sh;
}

void C4_Model_add_portal(C4_Model *m, const char *name, C4_Calc *calc, int aq, int at) {
    Portal *p = (Portal *)calloc(1, sizeof(*p));
    p->name = dup_str(name);
    p->calc = calc;
    p->advance_query = aq;
    p->advance_target = at;
    vec_push(&m->portals, p);
}

void C4_Model_add_span(C4_Model *m, const char *name, C4_State *state, int min_q, int max_q, int min_t,
                       int max_t) {
    Span *s = (Span *)calloc(1, sizeof(*s));
    s->name = dup_str(name);
    s->state = state;
    s->min_query = min_q; s->max_query = max_q; s->min_target = min_t; s->max_target = max_t;
    vec_push(&m->spans, s);
}

void C4_Model_configure_start_state(C4_Model *m, C4_Scope scope) { m->start_scope = scope; }
void C4_Model_configure_end_state(C4_Model *m, C4_Scope scope) { m->end_scope = scope; }
void C4_Model_open(C4_Model *m) { m->is_open = 1; }

int C4_Model_select_transitions(C4_Model *m, C4_Label label, C4_Transition **out, int max) {
    int i, n = 0;
    for (i = 0; i < m->transitions.n; i++) {
        C4_Transition *t = (C4_Transition *)m->transitions.v[i];
        if (t->label == label) {
            if (n < max) out[n] = t;
            n++;
        }
    }
    return n;
}
C4_Transition *C4_Model_select_single_transition(C4_Model *m, C4_Label label) {
    C4_Transition *t = NULL;
    return C4_Model_select_transitions(m, label, &t, 1) == 1 ? t : NULL;
}

/* ---- C4_Model_make_stereo (c4.c:681-770): duplicate everything but START/END */
void C4_Model_make_stereo(C4_Model *m, const char *suffix_a, const char *suffix_b) {
    const int ns = m->states.n, nt = m->transitions.n, nsh = m->shadows.n;
    C4_State **smap = (C4_State **)calloc((size_t)ns, sizeof(*smap));
    C4_Transition **tmap = (C4_Transition **)calloc((size_t)nt, sizeof(*tmap));
    int i, j;
    char *name;
    for (i = 0; i < ns; i++) {
        C4_State *s = (C4_State *)m->states.v[i];
        if (s != m->start && s != m->end) {
            name = cat_str(s->name, " ", suffix_b);
            smap[s->id] = C4_Model_add_state(m, name);
            free(name);
        }
    }
    for (i = 0; i < nt; i++) {
        C4_Transition *t = (C4_Transition *)m->transitions.v[i];
        name = cat_str(t->name, " ", suffix_b);
        /* START/END map to NULL, i.e. stay START/END */
        tmap[t->id] = C4_Model_add_transition(m, name, smap[t->input->id], smap[t->output->id],
                                              t->advance_query, t->advance_target, t->calc, t->label);
        free(name);
    }
    for (i = 0; i < nsh; i++) {
        C4_Shadow *sh = (C4_Shadow *)m->shadows.v[i], *nsw;
        name = cat_str(sh->name, " ", suffix_b);
        nsw = C4_Model_add_shadow(m, name, smap[((C4_State *)sh->src_states.v[0])->id],
                                  tmap[((C4_Transition *)sh->dst_transitions.v[0])->id], sh->start);
        free(name);
        for (j = 1; j < sh->src_states.n; j++) C4_Shadow_add_src_state(nsw, (C4_State *)sh->src_states.v[j]);
        for (j = 1; j < sh->dst_transitions.n; j++)
            C4_Shadow_add_dst_transition(nsw, (C4_Transition *)sh->dst_transitions.v[j]);
    }
    for (i = 0; i < ns; i++) {
        C4_State *s = (C4_State *)m->states.v[i];
        if (s != m->start && s != m->end) {
            name = cat_str(s->name, " ", suffix_a);
            free(s->name);
            s->name = name;
        }
    }
    for (i = 0; i < nt; i++) {
        C4_Transition *t = (C4_Transition *)m->transitions.v[i];
        name = cat_str(t->name, " ", suffix_a);
        free(t->name);
        t->name = name;
    }
    for (i = 0; i < nsh; i++) {
        C4_Shadow *sh = (C4_Shadow *)m->shadows.v[i];
        name = cat_str(sh->name, " ", suffix_a);
        free(sh->name);
        sh->name = name;
    }
    free(smap);
    free(tmap);
}

/* ---- C4_Model_insert (c4.c:772-998) ---------------------------------------- */
static int calc_same(const C4_Calc *a, const C4_Calc *b) {
    /* C4_Calc_diff (c4.c:87-95) compares max_score, the three callbacks and
     * protect; the device kind + params stand for the callbacks here */
    return a->max_score == b->max_score && a->dev.kind == b->dev.kind && a->dev.protect == b->dev.protect &&
           !memcmp(a->dev.param, b->dev.param, sizeof(a->dev.param));
}

void C4_Model_insert(C4_Model *target, C4_Model *insert, C4_State *src, C4_State *dst) {
    C4_Calc **cmap = (C4_Calc **)calloc((size_t)insert->calcs.n + 1, sizeof(*cmap));
    C4_State **smap = (C4_State **)calloc((size_t)insert->states.n, sizeof(*smap));
    C4_Transition **tmap = (C4_Transition **)calloc((size_t)insert->transitions.n + 1, sizeof(*tmap));
    int i, j;
    if (!src) src = target->start;
    if (!dst) dst = target->end;
    for (i = 0; i < insert->calcs.n; i++) { /* reuse an identical calc of the target */
        C4_Calc *ic = (C4_Calc *)insert->calcs.v[i], *tc = NULL;
        for (j = 0; j < target->calcs.n && !tc; j++)
            if (calc_same((C4_Calc *)target->calcs.v[j], ic)) tc = (C4_Calc *)target->calcs.v[j];
        if (!tc) tc = C4_Model_add_calc(target, ic->name, ic->max_score, ic->dev.kind, ic->dev.param,
                                        (C4_Protect)ic->dev.protect);
        cmap[ic->id] = tc;
    }
    for (i = 0; i < insert->states.n; i++) {
        C4_State *s = (C4_State *)insert->states.v[i];
        if (s != insert->start && s != insert->end) smap[s->id] = C4_Model_add_state(target, s->name);
    }
    smap[insert->start->id] = src;
    smap[insert->end->id] = dst;
    for (i = 0; i < insert->transitions.n; i++) {
        C4_Transition *t = (C4_Transition *)insert->transitions.v[i];
        tmap[t->id] = C4_Model_add_transition(target, t->name, smap[t->input->id], smap[t->output->id],
                                              t->advance_query, t->advance_target,
                                              t->calc ? cmap[t->calc->id] : NULL, t->label);
    }
    for (i = 0; i < insert->shadows.n; i++) {
        C4_Shadow *sh = (C4_Shadow *)insert->shadows.v[i];
        C4_Shadow *nsw = C4_Model_add_shadow(target, sh->name, smap[((C4_State *)sh->src_states.v[0])->id],
                                             tmap[((C4_Transition *)sh->dst_transitions.v[0])->id], sh->start);
        for (j = 1; j < sh->src_states.n; j++)
            C4_Shadow_add_src_state(nsw, smap[((C4_State *)sh->src_states.v[j])->id]);
        for (j = 1; j < sh->dst_transitions.n; j++)
            C4_Shadow_add_dst_transition(nsw, tmap[((C4_Transition *)sh->dst_transitions.v[j])->id]);
    }
    for (i = 0; i < insert->portals.n; i++) { /* c4.c:888-919: merge identical portals */
        Portal *ip = (Portal *)insert->portals.v[i];
        int found = 0;
        for (j = 0; j < target->portals.n && !found; j++) {
            Portal *tp = (Portal *)target->portals.v[j];
            found = tp->advance_query == ip->advance_query && tp->advance_target == ip->advance_target &&
                    calc_same(tp->calc, ip->calc);
        }
        if (!found) C4_Model_add_portal(target, ip->name, cmap[ip->calc->id], ip->advance_query, ip->advance_target);
    }
    for (i = 0; i < insert->spans.n; i++) {
        Span *sp = (Span *)insert->spans.v[i];
        C4_Model_add_span(target, sp->name, smap[sp->state->id], sp->min_query, sp->max_query,
                          sp->min_target, sp->max_target);
    }
    free(cmap);
    free(smap);
    free(tmap);
}

/* ---- closing (c4.c:1349-1383,1418-1486,1513-1680) --------------------------- */
static void set_ids(C4_Model *m) {
    int i;
    for (i = 0; i < m->states.n; i++) ((C4_State *)m->states.v[i])->id = i;
    for (i = 0; i < m->transitions.n; i++) ((C4_Transition *)m->transitions.v[i])->id = i;
    for (i = 0; i < m->shadows.n; i++) ((C4_Shadow *)m->shadows.v[i])->id = i;
    for (i = 0; i < m->calcs.n; i++) ((C4_Calc *)m->calcs.v[i])->id = i;
    for (i = 0; i < m->portals.n; i++) ((Portal *)m->portals.v[i])->id = i;
    for (i = 0; i < m->spans.n; i++) ((Span *)m->spans.v[i])->id = i;
}

static int is_silent(const C4_Transition *t) { return !t->advance_query && !t->advance_target; }

/* C4_Model_topological_sort (c4.c:1418-1486): silent transitions in dependency
 * order, then the emitting ones in insertion order, the whole list reversed.
 * The result is the per-cell evaluation order = the tie-break contract. */
static void topological_sort(C4_Model *m) {
    const int n = m->transitions.n;
    int *dependent = (int *)calloc((size_t)n, sizeof(int));
    C4_Transition **ordered = (C4_Transition **)malloc(sizeof(*ordered) * (size_t)n);
    int count = 0, i, j, removed;
    for (i = 0; i < n; i++) {
        C4_Transition *t = (C4_Transition *)m->transitions.v[i];
        if (!is_silent(t)) continue;
        for (j = 0; j < t->input->in_tr.n; j++) {
            C4_Transition *it = (C4_Transition *)t->input->in_tr.v[j];
            if (is_silent(it)) dependent[it->id]++;
        }
    }
    do {
        removed = 0;
        for (i = 0; i < n; i++) {
            C4_Transition *t = (C4_Transition *)m->transitions.v[i];
            if (dependent[i] != 0 || !is_silent(t)) continue;
            removed = 1;
            dependent[t->id] = -1;
            ordered[count++] = t;
            /* the reference decrements EVERY input transition of the input state
             * here (it tests the removed transition's own advances, c4.c:1457-1463) */
            for (j = 0; j < t->input->in_tr.n; j++) dependent[((C4_Transition *)t->input->in_tr.v[j])->id]--;
        }
    } while (removed);
    for (i = 0; i < n; i++) {
        C4_Transition *t = (C4_Transition *)m->transitions.v[i];
        if (!is_silent(t)) ordered[count++] = t;
    }
    if (count == n) {
        for (i = 0; i < n; i++) {
            C4_Transition *t = ordered[n - 1 - i];
            t->id = i;
            m->transitions.v[i] = t;
        }
    } else {
        set_err("model [%s] has a cycle of silent transitions", m->name);
    }
    free(ordered);
    free(dependent);
}

/* C4_Shadow_get_designation (c4.c:1539-1581): transitions a shadow's value may
 * travel along, found by walking input transitions back from its dst transitions. */
static void designate_recur(const C4_Shadow *sh, const C4_Transition *t, char *des, char *visited) {
    C4_State *s = t->input;
    int i;
    if (visited[s->id]) return;
    visited[s->id] = 1;
    for (i = 0; i < t->dst_shadows.n; i++)
        if (t->dst_shadows.v[i] == (void *)sh) return;
    for (i = 0; i < s->in_tr.n; i++) {
        C4_Transition *it = (C4_Transition *)s->in_tr.v[i];
        des[it->id] = 1;
        designate_recur(sh, it, des, visited);
    }
}

/* C4_Shadow_designation_fits (c4.c:1583-1624) */
static int designation_fits(const C4_Model *m, const char *a, const char *b) {
    const int nt = m->transitions.n, ns = m->states.n;
    char *used = (char *)calloc((size_t)ns, 1);
    int i, ok = 1;
    for (i = 0; i < nt && ok; i++)
        if (a[i] && b[i]) ok = 0;
    for (i = 0; i < nt; i++)
        if (a[i]) used[((C4_Transition *)m->transitions.v[i])->output->id] = 1;
    for (i = 0; i < nt && ok; i++)
        if (b[i] && used[((C4_Transition *)m->transitions.v[i])->input->id]) ok = 0;
    memset(used, 0, (size_t)ns);
    for (i = 0; i < nt; i++)
        if (b[i]) used[((C4_Transition *)m->transitions.v[i])->output->id] = 1;
    for (i = 0; i < nt && ok; i++)
        if (a[i] && used[((C4_Transition *)m->transitions.v[i])->input->id]) ok = 0;
    free(used);
    return ok;
}

/* C4_Model_designate_shadows (c4.c:1638-1667): first-fit packing of shadows into slots */
static void designate_shadows(C4_Model *m) {
    const int nt = m->transitions.n, ns = m->states.n;
    Vec slots = {0};
    int i, j;
    for (i = 0; i < m->shadows.n; i++) {
        C4_Shadow *sh = (C4_Shadow *)m->shadows.v[i];
        char *des = (char *)calloc((size_t)nt + 1, 1);
        char *visited = (char *)calloc((size_t)ns, 1);
        for (j = 0; j < sh->dst_transitions.n; j++) {
            C4_Transition *t = (C4_Transition *)sh->dst_transitions.v[j];
            des[t->id] = 1;
            designate_recur(sh, t, des, visited);
        }
        free(visited);
        sh->designation = -1;
        for (j = 0; j < slots.n && sh->designation < 0; j++) {
            char *master = (char *)slots.v[j];
            if (designation_fits(m, master, des)) {
                int k;
                for (k = 0; k < nt; k++) master[k] |= des[k];
                sh->designation = j;
            }
        }
        if (sh->designation < 0) {
            sh->designation = slots.n;
            vec_push(&slots, des);
        } else {
            free(des);
        }
    }
    m->total_shadow_designations = slots.n;
    for (i = 0; i < slots.n; i++) free(slots.v[i]);
    vec_free(&slots);
}

void C4_Model_close(C4_Model *m) {
    int i;
    host_error[0] = '\0';
    set_ids(m);
    topological_sort(m);
    designate_shadows(m);
    m->max_query_advance = m->max_target_advance = 0; /* C4_Model_finalise, c4.c:1513-1535 */
    for (i = 0; i < m->transitions.n; i++) {
        C4_Transition *t = (C4_Transition *)m->transitions.v[i];
        if (t->advance_query > m->max_query_advance) m->max_query_advance = t->advance_query;
        if (t->advance_target > m->max_target_advance) m->max_target_advance = t->advance_target;
    }
    m->is_open = 0;
}

/* ---- closed model -> engine tables ----------------------------------------- */
int C4_Model_flatten(const C4_Model *m, c4b_model *out) {
    int i, j;
    memset(out, 0, sizeof(*out));
    if (m->is_open) { set_err("model [%s] is still open", m->name); return -1; }
    if (host_error[0]) return -1;
    if (m->states.n > C4B_MAX_STATES || m->transitions.n > C4B_MAX_TRANSITIONS ||
        m->calcs.n > C4B_MAX_CALCS || m->total_shadow_designations > C4B_MAX_SHADOW_SLOTS) {
        set_err("model [%s] exceeds the engine's table sizes", m->name);
        return -1;
    }
    out->n_states = m->states.n;
    out->n_transitions = m->transitions.n;
    out->n_calcs = m->calcs.n;
    out->n_shadow_slots = m->total_shadow_designations;
    out->start_state = m->start->id;
    out->end_state = m->end->id;
    out->start_scope = (int32_t)m->start_scope;
    out->end_scope = (int32_t)m->end_scope;
    out->max_query_advance = m->max_query_advance;
    out->max_target_advance = m->max_target_advance;
    for (i = 0; i < m->calcs.n; i++) out->calcs[i] = ((C4_Calc *)m->calcs.v[i])->dev;
    for (i = 0; i < m->transitions.n; i++) {
        const C4_Transition *t = (const C4_Transition *)m->transitions.v[i];
        c4b_transition *o = &out->transitions[i];
        o->input = t->input->id;
        o->output = t->output->id;
        o->advance_query = t->advance_query;
        o->advance_target = t->advance_target;
        o->calc = t->calc ? t->calc->id : -1;
        o->label = (int32_t)t->label;
        /* a calc that reads a shadow value reads the slot of the shadow ENDING on
         * this transition (Viterbi_Row_shadow_end, viterbi.c:426-443) */
        if (t->calc && t->calc->dev.kind >= C4B_CALC_SPLICE_POST) {
            if (t->dst_shadows.n != 1) {
                set_err("transition [%s] needs exactly one ending shadow", t->name);
                return -1;
            }
            out->calcs[t->calc->id].param[2] = ((C4_Shadow *)t->dst_shadows.v[0])->designation;
        }
    }
    for (i = 0; i < m->shadows.n; i++) {
        const C4_Shadow *sh = (const C4_Shadow *)m->shadows.v[i];
        for (j = 0; j < sh->src_states.n; j++)
            out->shadow_start[((C4_State *)sh->src_states.v[j])->id][sh->designation] = (uint8_t)sh->start;
    }
    return 0;
}

/* ---- text dump (same record layout as oracle/ref_driver.c) ------------------ */
typedef struct {
    char *s;
    size_t n, cap;
} Str;
static void str_printf(Str *b, const char *fmt, ...) {
    va_list ap;
    int need;
    va_start(ap, fmt);
    need = vsnprintf(NULL, 0, fmt, ap);
    va_end(ap);
    if (b->n + (size_t)need + 1 > b->cap) {
        b->cap = (b->n + (size_t)need + 1) * 2;
        b->s = (char *)realloc(b->s, b->cap);
    }
    va_start(ap, fmt);
    vsnprintf(b->s + b->n, (size_t)need + 1, fmt, ap);
    va_end(ap);
    b->n += (size_t)need;
}

char *C4_Model_describe(const C4_Model *m) {
    Str b = {0};
    int i, j;
    str_printf(&b, "model name=\"%s\" states=%d transitions=%d calcs=%d shadows=%d portals=%d spans=%d"
                   " max_query_advance=%d max_target_advance=%d shadow_designations=%d"
                   " start_state=%d start_scope=%d end_state=%d end_scope=%d start_cell_func=0 end_cell_func=0\n",
               m->name, m->states.n, m->transitions.n, m->calcs.n, m->shadows.n, m->portals.n, m->spans.n,
               m->max_query_advance, m->max_target_advance, m->total_shadow_designations, m->start->id,
               (int)m->start_scope, m->end->id, (int)m->end_scope);
    for (i = 0; i < m->states.n; i++) {
        const C4_State *s = (const C4_State *)m->states.v[i];
        str_printf(&b, "state id=%d name=\"%s\" src_shadows=", s->id, s->name);
        for (j = 0; j < s->src_shadows.n; j++)
            str_printf(&b, "%s%d", j ? "," : "", ((C4_Shadow *)s->src_shadows.v[j])->id);
        str_printf(&b, "\n");
    }
    for (i = 0; i < m->calcs.n; i++) {
        const C4_Calc *c = (const C4_Calc *)m->calcs.v[i];
        str_printf(&b, "calc id=%d name=\"%s\" max_score=%d protect=%d kind=%d param=%d,%d,%d,%d\n", c->id,
                   c->name, c->max_score, c->dev.protect, c->dev.kind, c->dev.param[0], c->dev.param[1],
                   c->dev.param[2], c->dev.param[3]);
    }
    for (i = 0; i < m->transitions.n; i++) {
        const C4_Transition *t = (const C4_Transition *)m->transitions.v[i];
        str_printf(&b, "transition id=%d name=\"%s\" input=%d output=%d advance_query=%d advance_target=%d"
                       " calc=%d label=%d dst_shadows=",
                   t->id, t->name, t->input->id, t->output->id, t->advance_query, t->advance_target,
                   t->calc ? t->calc->id : -1, (int)t->label);
        for (j = 0; j < t->dst_shadows.n; j++)
            str_printf(&b, "%s%d", j ? "," : "", ((C4_Shadow *)t->dst_shadows.v[j])->id);
        str_printf(&b, "\n");
    }
    for (i = 0; i < m->shadows.n; i++) {
        const C4_Shadow *sh = (const C4_Shadow *)m->shadows.v[i];
        str_printf(&b, "shadow id=%d name=\"%s\" designation=%d src_states=", sh->id, sh->name, sh->designation);
        for (j = 0; j < sh->src_states.n; j++)
            str_printf(&b, "%s%d", j ? "," : "", ((C4_State *)sh->src_states.v[j])->id);
        str_printf(&b, " dst_transitions=");
        for (j = 0; j < sh->dst_transitions.n; j++)
            str_printf(&b, "%s%d", j ? "," : "", ((C4_Transition *)sh->dst_transitions.v[j])->id);
        str_printf(&b, " start=%d\n", (int)sh->start);
    }
    for (i = 0; i < m->portals.n; i++) {
        const Portal *p = (const Portal *)m->portals.v[i];
        str_printf(&b, "portal id=%d name=\"%s\" advance_query=%d advance_target=%d calc=%d\n", p->id, p->name,
                   p->advance_query, p->advance_target, p->calc ? p->calc->id : -1);
    }
    for (i = 0; i < m->spans.n; i++) {
        const Span *sp = (const Span *)m->spans.v[i];
        str_printf(&b, "span id=%d name=\"%s\" state=%d min_query=%d max_query=%d min_target=%d max_target=%d\n",
                   sp->id, sp->name, sp->state->id, sp->min_query, sp->max_query, sp->min_target, sp->max_target);
    }
    return b.s;
}
