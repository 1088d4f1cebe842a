// affine_traceback.cuh -- band planning, traceback walk and result assembly
// for the affine systolic path.
//
// Replaces Viterbi_Data_create_Alignment (src/c4/viterbi.c:342-392): follow the
// recorded winning transition from the END cell back to the transition that
// left START, then emit the path forwards, run-length merged exactly like
// Alignment_add (src/c4/alignment.c:75-100).
#pragma once
#include "c4b_common.cuh"

namespace c4b {

// What the traceback needs to know about one lattice.
struct TbJob {
    int32_t pair;       // index into the lattice array that was filled with TB
    int32_t result;     // index of the caller-visible result
    int32_t q_origin;   // sequence coordinate of lattice row 0
    int32_t t_origin;   // sequence coordinate of lattice column 0 (before banding)
    int32_t expect;     // 1: (score,end) must equal the score-only pass (banded)
    int32_t score_slot; // slot of the score-only pass result, or -1
    int64_t ops_off;    // first (transition,length) pair of this job's ops slot
    int32_t ops_cap;    // capacity of the slot in pairs
    int32_t reserved;   // record format: 1 = tag format (both traceback fills write it); 0 = decoded layout
};

// Longest target span an optimal local path ending at lattice row end_i can
// have: every prefix of the path scores > 0, each match adds <= max_sub and
// each deleted target position costs >= gap_min (DESIGN.md "band bound").
__host__ __device__ inline int64_t affine_band_width(int64_t rows, int max_sub, int gap_min) {
    if (max_sub <= 0) return 0;
    if (gap_min <= 0) return INT64_MAX / 4;
    return rows + (rows * (int64_t)max_sub) / gap_min + 1;
}

// After the score-only pass: restrict the traceback fill to the band of
// columns that can contain the optimal path, rows 0..end_i.
__global__ void affine_plan_band_kernel(const AffPair *__restrict__ full, const AffOut *__restrict__ score,
                                        AffPair *__restrict__ band, int32_t *__restrict__ band_j0,
                                        int n, int max_sub, int gap_min) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    AffPair P = full[p];
    const AffOut o = score[P.out_index];
    const int64_t W = affine_band_width(o.end_i, max_sub, gap_min);
    int64_t j0 = (int64_t)o.end_j - W;
    if (j0 < 0) j0 = 0;
    P.t = P.t + j0;
    P.blk_j0 = (int32_t)j0;   // SubOpt lists stay in full-lattice columns
    P.T = o.end_j - (int32_t)j0;
    P.Q = o.end_i;
    band[p] = P;
    band_j0[p] = (int32_t)j0;
}

template <int R>
__device__ __forceinline__ uint32_t tb_nibble(const uint32_t *__restrict__ tb, int nsteps, int i, int j) {
    constexpr int WPL = R / 8;
    const int sweep = i / (32 * R);
    const int ln = (i / R) & 31;
    const int r = i % R;
    const size_t word = (((size_t)sweep * nsteps + (j + ln)) * 32 + ln) * WPL + (r >> 3);
    return (tb[word] >> (4 * (r & 7))) & 15u;
}

// One thread per lattice: pointer-chasing walk, O(path length).
template <int R>
__global__ void affine_traceback_kernel(const AffPair *__restrict__ pairs, const AffOut *__restrict__ outs,
                                        const AffOut *__restrict__ score_outs,
                                        const int32_t *__restrict__ band_j0,
                                        const TbJob *__restrict__ jobs, int n, const AffModel mdl,
                                        c4b_result *__restrict__ results, int32_t *__restrict__ ops) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const TbJob J = jobs[g];
    const AffPair P = pairs[J.pair];
    const AffOut o = outs[P.out_index];
    const int nsteps = P.T + 1 + 31;
    const int j0 = band_j0 ? band_j0[J.pair] : 0;
    c4b_result res;
    res.score = o.best;
    res.status = 0;
    res.reserved = 0;
    res.ops_offset = J.ops_off;
    res.n_ops = 0;
    if (J.expect) {
        const AffOut so = score_outs[J.score_slot];
        // the banded refill must reproduce the full-lattice optimum at its corner
        if (so.best != o.best || o.end_i != P.Q || o.end_j != P.T) res.status = 3;
        res.score = so.best;
    }
    int i = o.end_i, j = o.end_j;
    res.query_end = J.q_origin + i;
    res.target_end = J.t_origin + j0 + j;
    int32_t *out = ops + 2 * J.ops_off;
    int n_runs = 0;
    int last_t = -1;
    bool overflow = false;
    auto emit = [&](int t) {
        if (t == last_t) {
            out[2 * (n_runs - 1) + 1] += 1;
        } else if (n_runs < J.ops_cap) {
            out[2 * n_runs] = t;
            out[2 * n_runs + 1] = 1;
            ++n_runs;
            last_t = t;
        } else {
            overflow = true;
        }
    };
    if (res.status == 0) {
        int state = 0;  // 0 = match, 1 = delete, 2 = insert
        emit(mdl.tME);
        for (;;) {
            uint32_t nib = tb_nibble<R>(P.tb, nsteps, i, j);
            // tag format of affine_fill16tb_kernel: bits 0-1 rank of the M winner (3 match ..
            // 0 insert), bit 2 D extended, bit 3 I extended -> this kernel's own layout
            if (J.reserved) nib = (((nib & 3u) ^ 3u) << 2) | ((nib & 4u) ? 0u : 2u) | ((nib & 8u) ? 0u : 1u);
            if (state == 0) {
                const uint32_t dir = nib >> 2;
                if (dir == 0) { emit(mdl.tMM); --i; --j; }
                else if (dir == 1) { emit(mdl.tSM); break; }
                else if (dir == 2) { emit(mdl.tDM); state = 1; }
                else { emit(mdl.tIM); state = 2; }
            } else if (state == 1) {
                if (nib & 2u) { emit(mdl.tMD); state = 0; }
                else emit(mdl.tDD);
                --j;
            } else {
                if (nib & 1u) { emit(mdl.tMI); state = 0; }
                else emit(mdl.tII);
                --i;
            }
            if (i < 0 || j < 0 || overflow) { res.status = 4; break; }
        }
        // runs were collected END -> START; flip to path order
        for (int a = 0, b = n_runs - 1; a < b; ++a, --b) {
            const int t0 = out[2 * a], l0 = out[2 * a + 1];
            out[2 * a] = out[2 * b];
            out[2 * a + 1] = out[2 * b + 1];
            out[2 * b] = t0;
            out[2 * b + 1] = l0;
        }
    }
    res.n_ops = n_runs;
    res.query_start = J.q_origin + (i < 0 ? 0 : i);
    res.target_start = J.t_origin + j0 + (j < 0 ? 0 : j);
    results[J.result] = res;
}

// Score-only batches: turn the fill's AffOut into caller results.
__global__ void affine_score_results_kernel(const AffPair *__restrict__ pairs, const AffOut *__restrict__ outs,
                                            const int32_t *__restrict__ q_origin,
                                            const int32_t *__restrict__ t_origin, int n,
                                            c4b_result *__restrict__ results) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const AffOut o = outs[pairs[p].out_index];
    c4b_result r;
    r.score = o.best;
    r.query_start = q_origin[p];
    r.target_start = t_origin[p];
    r.query_end = q_origin[p] + o.end_i;
    r.target_end = t_origin[p] + o.end_j;
    r.n_ops = 0;
    r.ops_offset = 0;
    r.status = 0;
    r.reserved = 0;
    results[pairs[p].out_index] = r;
}

}  // namespace c4b
