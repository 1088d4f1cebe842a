/* glib.c -- minimal glib-2 compatible shim (TEST INFRASTRUCTURE ONLY).
 * See glib.h.  Written from the public glib API documentation; only the
 * behaviour the reference relies on is provided.
 */
#include "glib.h"
#include <ctype.h>
#include <strings.h>

/* ---- memory ---- */
static void g_shim_oom(gsize n) {
    fprintf(stderr, "glibshim: out of memory allocating %lu bytes\n", n);
    abort();
}
gpointer g_malloc(gsize n) {
    gpointer p;
    if (!n) return NULL;
    p = malloc(n);
    if (!p) g_shim_oom(n);
    return p;
}
gpointer g_malloc0(gsize n) {
    gpointer p;
    if (!n) return NULL;
    p = calloc(1, n);
    if (!p) g_shim_oom(n);
    return p;
}
gpointer g_realloc(gpointer p, gsize n) {
    gpointer q;
    if (!n) {
        free(p);
        return NULL;
    }
    q = realloc(p, n);
    if (!q) g_shim_oom(n);
    return q;
}
void g_free(gpointer p) { free(p); }

/* ---- logging ---- */
static GLogFunc shim_handler = NULL;
static GLogLevelFlags shim_handler_levels = (GLogLevelFlags)0;
static gpointer shim_handler_data = NULL;

guint g_log_set_handler(const gchar *log_domain, GLogLevelFlags log_levels,
                        GLogFunc log_func, gpointer user_data) {
    (void)log_domain;
    shim_handler = log_func;
    shim_handler_levels = log_levels;
    shim_handler_data = user_data;
    return 1;
}

void g_shim_log(GLogLevelFlags level, const gchar *fmt, ...) {
    va_list ap;
    gchar *msg;
    va_start(ap, fmt);
    msg = g_strdup_vprintf(fmt, ap);
    va_end(ap);
    if (shim_handler && (shim_handler_levels & level)) {
        shim_handler(NULL, level, msg, shim_handler_data);
    } else {
        const char *tag = "Message";
        if (level & G_LOG_LEVEL_ERROR) tag = "ERROR";
        else if (level & G_LOG_LEVEL_CRITICAL) tag = "CRITICAL";
        else if (level & G_LOG_LEVEL_WARNING) tag = "WARNING";
        fflush(stdout);
        fprintf(stderr, "%s%s: %s\n", (level & G_LOG_LEVEL_MESSAGE) ? "" : "** ",
                tag, msg);
        fflush(stderr);
    }
    g_free(msg);
}

void g_print(const gchar *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stdout, fmt, ap);
    va_end(ap);
}

void g_on_error_stack_trace(const gchar *prg_name) { (void)prg_name; }

/* ---- strings ---- */
gchar *g_strdup(const gchar *s) {
    gchar *r;
    gsize n;
    if (!s) return NULL;
    n = strlen(s) + 1;
    r = (gchar *)g_malloc(n);
    memcpy(r, s, n);
    return r;
}
gchar *g_strndup(const gchar *s, gsize n) {
    gchar *r;
    if (!s) return NULL;
    r = (gchar *)g_malloc0(n + 1);
    strncpy(r, s, n);
    r[n] = '\0';
    return r;
}
gchar *g_strnfill(gsize length, gchar fill_char) {
    gchar *r = (gchar *)g_malloc(length + 1);
    memset(r, fill_char, length);
    r[length] = '\0';
    return r;
}
gchar *g_strdup_vprintf(const gchar *fmt, va_list args) {
    va_list cp;
    int n;
    gchar *r;
    va_copy(cp, args);
    n = vsnprintf(NULL, 0, fmt, cp);
    va_end(cp);
    if (n < 0) n = 0;
    r = (gchar *)g_malloc((gsize)n + 1);
    vsnprintf(r, (size_t)n + 1, fmt, args);
    return r;
}
gchar *g_strdup_printf(const gchar *fmt, ...) {
    va_list ap;
    gchar *r;
    va_start(ap, fmt);
    r = g_strdup_vprintf(fmt, ap);
    va_end(ap);
    return r;
}
gchar *g_strconcat(const gchar *first, ...) {
    va_list ap;
    gsize total;
    const gchar *s;
    gchar *r, *p;
    if (!first) return NULL;
    total = strlen(first);
    va_start(ap, first);
    while ((s = va_arg(ap, const gchar *))) total += strlen(s);
    va_end(ap);
    r = (gchar *)g_malloc(total + 1);
    p = r;
    strcpy(p, first);
    p += strlen(first);
    va_start(ap, first);
    while ((s = va_arg(ap, const gchar *))) {
        strcpy(p, s);
        p += strlen(s);
    }
    va_end(ap);
    return r;
}
gchar **g_strsplit(const gchar *string, const gchar *delimiter, gint max_tokens) {
    GPtrArray *out = g_ptr_array_new();
    const gchar *rest = string, *hit;
    gsize dl = strlen(delimiter);
    gchar **res;
    if (max_tokens < 1) max_tokens = INT_MAX;
    if (*rest) {
        while (--max_tokens && (hit = strstr(rest, delimiter))) {
            g_ptr_array_add(out, g_strndup(rest, (gsize)(hit - rest)));
            rest = hit + dl;
        }
        g_ptr_array_add(out, g_strdup(rest));
    }
    g_ptr_array_add(out, NULL);
    res = (gchar **)g_ptr_array_free(out, FALSE);
    return res;
}
gchar *g_strjoinv(const gchar *separator, gchar **str_array) {
    GString *s = g_string_new("");
    gint i;
    if (!separator) separator = "";
    for (i = 0; str_array[i]; i++) {
        if (i) g_string_append(s, separator);
        g_string_append(s, str_array[i]);
    }
    return g_string_free(s, FALSE);
}
void g_strfreev(gchar **str_array) {
    gint i;
    if (!str_array) return;
    for (i = 0; str_array[i]; i++) g_free(str_array[i]);
    g_free(str_array);
}
gchar *g_strchug(gchar *string) {
    gchar *start = string;
    while (*start && isspace((unsigned char)*start)) start++;
    memmove(string, start, strlen(start) + 1);
    return string;
}
gchar *g_strchomp(gchar *string) {
    gsize len = strlen(string);
    while (len && isspace((unsigned char)string[len - 1])) string[--len] = '\0';
    return string;
}
gint g_strcasecmp(const gchar *s1, const gchar *s2) { return strcasecmp(s1, s2); }
gchar *g_strup(gchar *string) {
    gchar *p;
    for (p = string; *p; p++) *p = (gchar)toupper((unsigned char)*p);
    return string;
}
const gchar *g_getenv(const gchar *variable) { return getenv(variable); }

/* ---- GString ---- */
static void g_string_reserve(GString *s, gsize need) {
    if (need + 1 > s->allocated_len) {
        gsize n = s->allocated_len ? s->allocated_len : 16;
        while (n < need + 1) n <<= 1;
        s->str = (gchar *)g_realloc(s->str, n);
        s->allocated_len = n;
    }
}
GString *g_string_sized_new(gsize dfl_size) {
    GString *s = g_new0(GString, 1);
    g_string_reserve(s, dfl_size > 2 ? dfl_size : 2);
    s->str[0] = '\0';
    return s;
}
GString *g_string_new(const gchar *init) {
    GString *s = g_string_sized_new(init ? strlen(init) + 2 : 2);
    if (init) g_string_append(s, init);
    return s;
}
GString *g_string_append(GString *s, const gchar *val) {
    gsize n = strlen(val);
    g_string_reserve(s, s->len + n);
    memcpy(s->str + s->len, val, n + 1);
    s->len += n;
    return s;
}
GString *g_string_append_c(GString *s, gchar c) {
    g_string_reserve(s, s->len + 1);
    s->str[s->len++] = c;
    s->str[s->len] = '\0';
    return s;
}
GString *g_string_truncate(GString *s, gsize len) {
    if (len < s->len) s->len = len;
    s->str[s->len] = '\0';
    return s;
}
gchar *g_string_free(GString *s, gboolean free_segment) {
    gchar *seg = s->str;
    if (free_segment) {
        g_free(seg);
        seg = NULL;
    }
    g_free(s);
    return seg;
}

/* ---- GStringChunk: every string separately allocated ---- */
struct _GStringChunk {
    GPtrArray *strings;
};
GStringChunk *g_string_chunk_new(gsize size) {
    GStringChunk *c = g_new0(GStringChunk, 1);
    (void)size;
    c->strings = g_ptr_array_new();
    return c;
}
gchar *g_string_chunk_insert(GStringChunk *chunk, const gchar *string) {
    gchar *s = g_strdup(string);
    g_ptr_array_add(chunk->strings, s);
    return s;
}
void g_string_chunk_free(GStringChunk *chunk) {
    guint i;
    for (i = 0; i < chunk->strings->len; i++) g_free(chunk->strings->pdata[i]);
    g_ptr_array_free(chunk->strings, TRUE);
    g_free(chunk);
}

/* ---- GPtrArray ---- */
typedef struct {
    gpointer *pdata;
    guint len;
    guint alloc;
} RealPtrArray;
static void ptr_array_reserve(RealPtrArray *a, guint need) {
    if (need > a->alloc) {
        guint n = a->alloc ? a->alloc : 16;
        while (n < need) n <<= 1;
        a->pdata = (gpointer *)g_realloc(a->pdata, sizeof(gpointer) * n);
        memset(a->pdata + a->alloc, 0, sizeof(gpointer) * (n - a->alloc));
        a->alloc = n;
    }
}
GPtrArray *g_ptr_array_new(void) {
    RealPtrArray *a = g_new0(RealPtrArray, 1);
    return (GPtrArray *)a;
}
void g_ptr_array_add(GPtrArray *array, gpointer data) {
    RealPtrArray *a = (RealPtrArray *)array;
    ptr_array_reserve(a, a->len + 1);
    a->pdata[a->len++] = data;
}
gpointer *g_ptr_array_free(GPtrArray *array, gboolean free_seg) {
    gpointer *seg = array->pdata;
    if (free_seg) {
        g_free(seg);
        seg = NULL;
    }
    g_free(array);
    return seg;
}
void g_ptr_array_set_size(GPtrArray *array, gint length) {
    RealPtrArray *a = (RealPtrArray *)array;
    guint n = (guint)length;
    if (n > a->len) {
        guint i;
        ptr_array_reserve(a, n);
        for (i = a->len; i < n; i++) a->pdata[i] = NULL;
    }
    a->len = n;
}
gboolean g_ptr_array_remove_fast(GPtrArray *array, gpointer data) {
    guint i;
    for (i = 0; i < array->len; i++) {
        if (array->pdata[i] == data) {
            array->pdata[i] = array->pdata[array->len - 1];
            array->pdata[--array->len] = NULL;
            return TRUE;
        }
    }
    return FALSE;
}

/* ---- GArray ---- */
typedef struct {
    gchar *data;
    guint len;
    guint alloc; /* in elements */
    guint elt_size;
    gboolean zero_terminated;
    gboolean clear;
} RealArray;
static void array_reserve(RealArray *a, guint need) {
    guint want = need + (a->zero_terminated ? 1 : 0);
    if (want > a->alloc) {
        guint n = a->alloc ? a->alloc : 16;
        while (n < want) n <<= 1;
        a->data = (gchar *)g_realloc(a->data, (gsize)n * a->elt_size);
        memset(a->data + (gsize)a->alloc * a->elt_size, 0,
               (gsize)(n - a->alloc) * a->elt_size);
        a->alloc = n;
    }
}
GArray *g_array_new(gboolean zero_terminated, gboolean clear_, guint element_size) {
    RealArray *a = g_new0(RealArray, 1);
    a->elt_size = element_size;
    a->zero_terminated = zero_terminated;
    a->clear = clear_;
    if (zero_terminated) array_reserve(a, 0);
    return (GArray *)a;
}
gchar *g_array_free(GArray *array, gboolean free_segment) {
    gchar *seg = array->data;
    if (free_segment) {
        g_free(seg);
        seg = NULL;
    }
    g_free(array);
    return seg;
}
GArray *g_array_append_vals(GArray *array, gconstpointer data, guint len) {
    RealArray *a = (RealArray *)array;
    array_reserve(a, a->len + len);
    memcpy(a->data + (gsize)a->len * a->elt_size, data, (gsize)len * a->elt_size);
    a->len += len;
    if (a->zero_terminated)
        memset(a->data + (gsize)a->len * a->elt_size, 0, a->elt_size);
    return array;
}
GArray *g_array_set_size(GArray *array, guint length) {
    RealArray *a = (RealArray *)array;
    if (length > a->len) {
        array_reserve(a, length);
        memset(a->data + (gsize)a->len * a->elt_size, 0,
               (gsize)(length - a->len) * a->elt_size);
    }
    a->len = length;
    if (a->zero_terminated && a->data)
        memset(a->data + (gsize)a->len * a->elt_size, 0, a->elt_size);
    return array;
}

/* ---- GTree: sorted array + binary search (reference trees are tiny) ---- */
struct _GTree {
    GCompareFunc cmp;
    gpointer *keys;
    gpointer *vals;
    guint len, alloc;
};
GTree *g_tree_new(GCompareFunc key_compare_func) {
    GTree *t = g_new0(GTree, 1);
    t->cmp = key_compare_func;
    return t;
}
static gboolean tree_find(GTree *t, gconstpointer key, guint *pos) {
    guint lo = 0, hi = t->len;
    while (lo < hi) {
        guint mid = lo + (hi - lo) / 2;
        gint c = t->cmp(key, t->keys[mid]);
        if (c == 0) {
            *pos = mid;
            return TRUE;
        }
        if (c < 0) hi = mid;
        else lo = mid + 1;
    }
    *pos = lo;
    return FALSE;
}
void g_tree_insert(GTree *t, gpointer key, gpointer value) {
    guint pos;
    if (tree_find(t, key, &pos)) {
        t->vals[pos] = value; /* glib: replaces value, keeps old key */
        return;
    }
    if (t->len + 1 > t->alloc) {
        guint n = t->alloc ? t->alloc * 2 : 16;
        t->keys = (gpointer *)g_realloc(t->keys, sizeof(gpointer) * n);
        t->vals = (gpointer *)g_realloc(t->vals, sizeof(gpointer) * n);
        t->alloc = n;
    }
    memmove(t->keys + pos + 1, t->keys + pos, sizeof(gpointer) * (t->len - pos));
    memmove(t->vals + pos + 1, t->vals + pos, sizeof(gpointer) * (t->len - pos));
    t->keys[pos] = key;
    t->vals[pos] = value;
    t->len++;
}
gpointer g_tree_lookup(GTree *t, gconstpointer key) {
    guint pos;
    if (tree_find(t, key, &pos)) return t->vals[pos];
    return NULL;
}
void g_tree_destroy(GTree *t) {
    g_free(t->keys);
    g_free(t->vals);
    g_free(t);
}
