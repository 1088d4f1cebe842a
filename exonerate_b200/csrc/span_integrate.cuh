// span_integrate.cuh -- BSDP span integration on the device.
//
// Heuristic_Span_integrate (src/bsdp/heuristic.c:589-678): for every cell (i,j) of the dst
// region, the src-region cell with the best score among those a span of
// [min_query,max_query] x [min_target,max_target] symbols can bridge -- the FIRST such cell in
// (query, target) scan order on ties (strict '<', :638) -- or (-1,-1) when the window is empty.
// Heuristic_Span_score is the constant 0 (:362-366), so the reference's reuse of the previous
// cell's answer when the window did not move (:618-621) is only a CPU shortcut.
//
// One thread per dst cell, a sequential scan of its window in the reference's order (the src
// matrix is a few thousand ints: L1-resident).  The windows of neighbouring cells overlap
// almost entirely, so a real batched version would share a running max per row; this first
// version exists to pin the semantics on the device for the batched SAR pass.
#pragma once
#include "c4b_common.cuh"

namespace c4b {

struct SpanArgs {
    int32_t sqs, sts, sql, stl;  // src region
    int32_t dqs, dts, dql, dtl;  // dst region
    int32_t min_q, max_q, min_t, max_t;
};

__global__ void span_integrate_kernel(const int32_t *__restrict__ src_scores, SpanArgs a,
                                      int32_t *__restrict__ positions) {
    const int cells = (a.dql + 1) * (a.dtl + 1);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += gridDim.x * blockDim.x) {
        const int i = c / (a.dtl + 1), j = c - i * (a.dtl + 1);
        const int iq = max(a.sqs, a.dqs + i - a.max_q), it = max(a.sts, a.dts + j - a.max_t);
        const int fq = min(a.sqs + a.sql, a.dqs + i - a.min_q), ft = min(a.sts + a.stl, a.dts + j - a.min_t);
        int top = LOW, top_q = -1, top_t = -1;
        for (int x = iq; x <= fq; ++x) {
            const long long base = (long long)(x - a.sqs) * (a.stl + 1) - a.sts;
            for (int y = it; y <= ft; ++y) {
                const int cand = __ldg(src_scores + (base + y));
                if (top < cand) { top = cand; top_q = x; top_t = y; }
            }
        }
        positions[2 * c] = top_q;
        positions[2 * c + 1] = top_t;
    }
}

// ---- the batched form (c4b_span_score_batch): integrate + START table in one pass -------------
// One span edge of BSDP (SAR_Span_find_score, src/bsdp/sar.c:898-917): the src fill left END's
// cell of every src cell in `src_end` (C_src ints per cell; cells END never reached still hold the
// fill pattern of the table, which reads as a score far below C4_IMPOSSIBLY_LOW_SCORE, i.e. what
// Heuristic_Span_clear leaves, heuristic.c:563-573).  For every dst cell: the best src cell its
// window reaches (Heuristic_Span_integrate) and, from it, START's cell for the dst fill exactly as
// Heuristic_Span_dst_init_start_func hands it out (heuristic.c:412-443): the src END cell, or the
// dummy cell {C4_IMPOSSIBLY_LOW_SCORE, 0, ...} when the window holds nothing reachable.
struct SpanJob {
    SpanArgs a;
    const int32_t *src_end;   // (sql+1) x (stl+1) x c_src
    int32_t *dst_start;       // (dql+1) x (dtl+1) x c_dst
    int32_t c_src, c_dst;
};

__global__ void span_start_table_kernel(const SpanJob *__restrict__ jobs) {
    const SpanJob J = jobs[blockIdx.y];
    const SpanArgs &a = J.a;
    const int cells = (a.dql + 1) * (a.dtl + 1);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += gridDim.x * blockDim.x) {
        const int i = c / (a.dtl + 1), j = c - i * (a.dtl + 1);
        const int iq = max(a.sqs, a.dqs + i - a.max_q), it = max(a.sts, a.dts + j - a.max_t);
        const int fq = min(a.sqs + a.sql, a.dqs + i - a.min_q), ft = min(a.sts + a.stl, a.dts + j - a.min_t);
        int top = LOW;
        long long top_cell = -1;
        for (int x = iq; x <= fq; ++x) {
            const long long base = (long long)(x - a.sqs) * (a.stl + 1) - a.sts;
            for (int y = it; y <= ft; ++y) {
                const int cand = J.src_end[(base + y) * J.c_src];
                if (top < cand) { top = cand; top_cell = base + y; }
            }
        }
        int32_t *out = J.dst_start + (size_t)c * J.c_dst;
        for (int l = 0; l < J.c_dst; ++l)
            out[l] = (top_cell < 0) ? (l == 0 ? LOW : 0) : (l < J.c_src ? J.src_end[top_cell * J.c_src + l] : 0);
    }
}

}  // namespace c4b
