/* splice.c -- per-position splice-site scores of a genomic sequence, the int32
 * arrays the SPLICE_PRE / SPLICE_POST calcs read (include/c4b200.h).
 *
 * Follows src/sequence/splice.c:66-120 (position frequency tables of Senapathy,
 * Shapiro & Harris, Methods Enzymol. 183:252-278), :241-300 (predictor set-up:
 * reverse-strand models, pseudocount, log-odds * 1.5 in float32), :320-344
 * (windowed sum, clipped at the sequence ends) and :379-397 (rounding to int).
 * The float32 accumulation order is kept so the integers come out identical;
 * tests/test_host_models.py compares with arrays exported from the reference.
 */
#include <math.h>
#include <string.h>

#include "c4host.h"

#define SPLICE_LOW (-987654321.0f)

typedef struct {
    int length, splice_after, gtag_only;
    char expect_one, expect_two;
    float data[16][5];
    unsigned char index[256];
} Predictor;

/* percent frequencies A C G T around the donor (9 positions, site after 3) */
static const int freq_5ss[9][4] = {{28, 40, 17, 14}, {59, 14, 13, 14}, {8, 5, 81, 6},   {0, 0, 100, 0},
                                   {0, 0, 0, 100},   {54, 2, 42, 2},   {74, 8, 11, 8}, {5, 6, 85, 4},
                                   {16, 18, 21, 45}};
/* ... and around the acceptor (15 positions, site after 14) */
static const int freq_3ss[15][4] = {{10, 31, 14, 44}, {8, 36, 14, 43}, {6, 34, 12, 48}, {6, 34, 8, 52},
                                    {9, 37, 9, 45},   {9, 38, 10, 44}, {8, 44, 9, 40},  {9, 41, 8, 41},
                                    {6, 44, 6, 45},   {6, 40, 6, 48},  {23, 28, 26, 23}, {2, 79, 1, 18},
                                    {100, 0, 0, 0},   {0, 0, 100, 0},  {28, 14, 47, 11}};

static void predictor_init(Predictor *sp, int type, int force_gtag) {
    const int is5 = (type == C4B_SPLICE_5_FORWARD || type == C4B_SPLICE_5_REVERSE);
    const int reverse = (type == C4B_SPLICE_5_REVERSE || type == C4B_SPLICE_3_REVERSE);
    int i, j, a, z;
    memset(sp, 0, sizeof(*sp));
    if (is5) {
        sp->length = 9;
        sp->splice_after = 3;
        for (i = 0; i < 9; i++)
            for (j = 0; j < 4; j++) sp->data[i][j] = (float)freq_5ss[i][j];
    } else {
        sp->length = 15;
        sp->splice_after = 14 - 2; /* the acceptor score sits on the AG, splice.c:214 */
        for (i = 0; i < 15; i++)
            for (j = 0; j < 4; j++) sp->data[i][j] = (float)freq_3ss[i][j];
    }
    if (reverse) { /* read the other strand: positions mirrored, bases complemented */
        for (a = 0, z = sp->length - 1; a < z; a++, z--)
            for (j = 0; j < 4; j++) {
                const float swap = sp->data[a][j];
                sp->data[a][j] = sp->data[z][j];
                sp->data[z][j] = swap;
            }
        sp->splice_after = sp->length - sp->splice_after - 2;
    }
    memset(sp->index, 4, sizeof(sp->index));
    if (!reverse) {
        sp->index['A'] = sp->index['a'] = 0; sp->index['C'] = sp->index['c'] = 1;
        sp->index['G'] = sp->index['g'] = 2; sp->index['T'] = sp->index['t'] = 3;
    } else {
        sp->index['T'] = sp->index['t'] = 0; sp->index['G'] = sp->index['g'] = 1;
        sp->index['C'] = sp->index['c'] = 2; sp->index['A'] = sp->index['a'] = 3;
    }
    for (i = 0; i < sp->length; i++) {
        for (j = 0; j < 4; j++) {
            /* pseudocount 1 over 25+1, then 1.5 * ln, each step stored as float32 */
            sp->data[i][j] = (float)(((float)(1 + sp->data[i][j])) / (25.0 + 1.0));
            sp->data[i][j] = (float)(log(sp->data[i][j]) * 1.5);
        }
        sp->data[i][4] = 0.0f;
    }
    sp->gtag_only = force_gtag;
    switch (type) {
    case C4B_SPLICE_5_FORWARD: sp->expect_one = 'G'; sp->expect_two = 'T'; break;
    case C4B_SPLICE_3_FORWARD: sp->expect_one = 'A'; sp->expect_two = 'G'; break;
    case C4B_SPLICE_5_REVERSE: sp->expect_one = 'A'; sp->expect_two = 'C'; break;
    default: sp->expect_one = 'C'; sp->expect_two = 'T'; break;
    }
}

static int to_upper(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

static float predict_position(const Predictor *sp, const uint8_t *seq, int seq_len, int pos) {
    float score = 0.0f;
    int seq_start = pos - sp->splice_after, model_start = 0, calc_length = sp->length, i;
    if (seq_start < 0) {
        model_start = -seq_start;
        seq_start = 0;
        calc_length -= model_start;
    }
    if (seq_start + calc_length > seq_len) calc_length = seq_len - seq_start;
    for (i = 0; i < calc_length; i++) score += sp->data[model_start + i][sp->index[seq[seq_start + i]]];
    if (sp->gtag_only &&
        !(pos + 1 < seq_len + 1 && to_upper(seq[pos]) == sp->expect_one &&
          to_upper(pos + 1 < seq_len ? seq[pos + 1] : 0) == sp->expect_two))
        return SPLICE_LOW;
    return score;
}

/* SplicePredictor_predict_array_int over a whole sequence for one site type */
void c4b_host_splice_array(int type, const uint8_t *seq, int32_t len, int force_gtag, int32_t *out) {
    Predictor sp;
    int i;
    predictor_init(&sp, type, force_gtag);
    for (i = 0; i < len; i++) {
        const float s = predict_position(&sp, seq, len, i);
        out[i] = (int32_t)((s < 0) ? (s - 0.5) : (s + 0.5));
    }
}

/* all four arrays, in C4B_SPLICE_* order, into out[4*len] */
void c4b_host_splice_arrays(const uint8_t *seq, int32_t len, int force_gtag, int32_t *out) {
    int t;
    for (t = 0; t < C4B_SPLICE_TOTAL; t++) c4b_host_splice_array(t, seq, len, force_gtag, out + (size_t)t * len);
}
