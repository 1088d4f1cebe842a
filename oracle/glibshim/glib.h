/* glib.h -- minimal glib-2 compatible shim (TEST INFRASTRUCTURE ONLY).
 *
 * glib is not installed in this image.  The unmodified reference
 * (nathanweeks/exonerate, C89 + glib) only uses ~55 glib entry points
 * (containers, strings, logging; no arithmetic), listed in SURVEY.md §8c.
 * This header + glib.c provide exactly those so that the reference sources
 * can be compiled WHERE THEY LIE under /root/reference into oracle/_ref/.
 * Nothing under exonerate_b200/ (the product) includes this file.
 */
#ifndef C4B_GLIB_SHIM_H
#define C4B_GLIB_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <stddef.h>
#include <limits.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLIB_MAJOR_VERSION 2
#define GLIB_MINOR_VERSION 0
#define GLIB_MICRO_VERSION 0

typedef char gchar;
typedef short gshort;
typedef long glong;
typedef int gint;
typedef gint gboolean;
typedef unsigned char guchar;
typedef unsigned short gushort;
typedef unsigned long gulong;
typedef unsigned int guint;
typedef float gfloat;
typedef double gdouble;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef signed char gint8;
typedef unsigned char guint8;
typedef short gint16;
typedef unsigned short guint16;
typedef int gint32;
typedef unsigned int guint32;
typedef long gint64;
typedef unsigned long guint64;
typedef unsigned long gsize;
typedef long gssize;

#define G_HAVE_GINT64 1
#define G_GINT64_CONSTANT(val) (val##L)
#define G_GNUC_EXTENSION __extension__
#define G_DIR_SEPARATOR '/'
#define G_DIR_SEPARATOR_S "/"

#define GUINT64_TO_BE(v) (__builtin_bswap64((guint64)(v)))
#define GUINT64_FROM_BE(v) (__builtin_bswap64((guint64)(v)))

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#ifndef NULL
#define NULL ((void *)0)
#endif
#undef MIN
#undef MAX
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#define ABS(a) (((a) < 0) ? -(a) : (a))
#define CLAMP(x, lo, hi) (((x) > (hi)) ? (hi) : (((x) < (lo)) ? (lo) : (x)))

#define GINT_TO_POINTER(i) ((gpointer)(glong)(i))
#define GPOINTER_TO_INT(p) ((gint)(glong)(p))
#define GUINT_TO_POINTER(u) ((gpointer)(gulong)(u))
#define GPOINTER_TO_UINT(p) ((guint)(gulong)(p))

typedef gint (*GCompareFunc)(gconstpointer a, gconstpointer b);

/* ---- memory ---- */
gpointer g_malloc(gsize n);
gpointer g_malloc0(gsize n);
gpointer g_realloc(gpointer p, gsize n);
void g_free(gpointer p);
#define g_new(type, n) ((type *)g_malloc(sizeof(type) * (gsize)(n)))
#define g_new0(type, n) ((type *)g_malloc0(sizeof(type) * (gsize)(n)))
#define g_renew(type, mem, n) ((type *)g_realloc((mem), sizeof(type) * (gsize)(n)))

/* ---- logging ---- */
typedef enum {
    G_LOG_FLAG_RECURSION = 1 << 0,
    G_LOG_FLAG_FATAL = 1 << 1,
    G_LOG_LEVEL_ERROR = 1 << 2,
    G_LOG_LEVEL_CRITICAL = 1 << 3,
    G_LOG_LEVEL_WARNING = 1 << 4,
    G_LOG_LEVEL_MESSAGE = 1 << 5,
    G_LOG_LEVEL_INFO = 1 << 6,
    G_LOG_LEVEL_DEBUG = 1 << 7
} GLogLevelFlags;

typedef void (*GLogFunc)(const gchar *log_domain, GLogLevelFlags log_level,
                         const gchar *message, gpointer user_data);
guint g_log_set_handler(const gchar *log_domain, GLogLevelFlags log_levels,
                        GLogFunc log_func, gpointer user_data);
void g_shim_log(GLogLevelFlags level, const gchar *fmt, ...)
    __attribute__((format(printf, 2, 3)));
void g_print(const gchar *fmt, ...) __attribute__((format(printf, 1, 2)));
void g_on_error_stack_trace(const gchar *prg_name);

#define g_error(...)                                  \
    do {                                              \
        g_shim_log(G_LOG_LEVEL_ERROR, __VA_ARGS__);   \
        abort();                                      \
    } while (0)
#define g_message(...) g_shim_log(G_LOG_LEVEL_MESSAGE, __VA_ARGS__)
#define g_warning(...) g_shim_log(G_LOG_LEVEL_WARNING, __VA_ARGS__)
#define g_critical(...) g_shim_log(G_LOG_LEVEL_CRITICAL, __VA_ARGS__)

#ifdef G_DISABLE_ASSERT
#define g_assert(expr) do { (void)0; } while (0)
#define g_assert_not_reached() do { (void)0; } while (0)
#else
#define g_assert(expr)                                                      \
    do {                                                                    \
        if (!(expr)) {                                                      \
            g_shim_log(G_LOG_LEVEL_ERROR,                                   \
                       "file %s: line %d (%s): assertion failed: (%s)",     \
                       __FILE__, __LINE__, __func__, #expr);                \
            abort();                                                        \
        }                                                                   \
    } while (0)
#define g_assert_not_reached()                                              \
    do {                                                                    \
        g_shim_log(G_LOG_LEVEL_ERROR, "file %s: line %d (%s): not reached", \
                   __FILE__, __LINE__, __func__);                           \
        abort();                                                            \
    } while (0)
#endif

/* ---- threads (reference never defines USE_PTHREADS) ---- */
#define g_thread_supported() (TRUE)
#define g_thread_init(x) do { (void)0; } while (0)

/* ---- strings ---- */
gchar *g_strdup(const gchar *s);
gchar *g_strndup(const gchar *s, gsize n);
gchar *g_strnfill(gsize length, gchar fill_char);
gchar *g_strdup_printf(const gchar *fmt, ...) __attribute__((format(printf, 1, 2)));
gchar *g_strdup_vprintf(const gchar *fmt, va_list args);
gchar *g_strconcat(const gchar *first, ...);
gchar **g_strsplit(const gchar *string, const gchar *delimiter, gint max_tokens);
gchar *g_strjoinv(const gchar *separator, gchar **str_array);
void g_strfreev(gchar **str_array);
gchar *g_strchug(gchar *string);
gchar *g_strchomp(gchar *string);
#define g_strstrip(string) g_strchomp(g_strchug(string))
gint g_strcasecmp(const gchar *s1, const gchar *s2);
gchar *g_strup(gchar *string);
const gchar *g_getenv(const gchar *variable);

typedef struct {
    gchar *str;
    gsize len;
    gsize allocated_len;
} GString;
GString *g_string_new(const gchar *init);
GString *g_string_sized_new(gsize dfl_size);
GString *g_string_append(GString *string, const gchar *val);
GString *g_string_append_c(GString *string, gchar c);
GString *g_string_truncate(GString *string, gsize len);
gchar *g_string_free(GString *string, gboolean free_segment);

typedef struct _GStringChunk GStringChunk;
GStringChunk *g_string_chunk_new(gsize size);
gchar *g_string_chunk_insert(GStringChunk *chunk, const gchar *string);
void g_string_chunk_free(GStringChunk *chunk);

/* ---- pointer arrays ---- */
typedef struct {
    gpointer *pdata;
    guint len;
} GPtrArray;
GPtrArray *g_ptr_array_new(void);
void g_ptr_array_add(GPtrArray *array, gpointer data);
gpointer *g_ptr_array_free(GPtrArray *array, gboolean free_seg);
void g_ptr_array_set_size(GPtrArray *array, gint length);
gboolean g_ptr_array_remove_fast(GPtrArray *array, gpointer data);
#define g_ptr_array_index(array, index_) ((array)->pdata)[index_]

/* ---- value arrays ---- */
typedef struct {
    gchar *data;
    guint len;
} GArray;
GArray *g_array_new(gboolean zero_terminated, gboolean clear_, guint element_size);
gchar *g_array_free(GArray *array, gboolean free_segment);
GArray *g_array_append_vals(GArray *array, gconstpointer data, guint len);
GArray *g_array_set_size(GArray *array, guint length);
#define g_array_append_val(a, v) g_array_append_vals(a, &(v), 1)
#define g_array_index(a, t, i) (((t *)(void *)(a)->data)[(i)])

/* ---- ordered map ---- */
typedef struct _GTree GTree;
GTree *g_tree_new(GCompareFunc key_compare_func);
void g_tree_insert(GTree *tree, gpointer key, gpointer value);
gpointer g_tree_lookup(GTree *tree, gconstpointer key);
void g_tree_destroy(GTree *tree);

typedef struct _GSList GSList;
struct _GSList {
    gpointer data;
    GSList *next;
};

#ifdef __cplusplus
}
#endif
#endif /* C4B_GLIB_SHIM_H */
