// hsp_extend.cuh -- batched HSP seeding / extension.
//
// Replaces the per-seed work of HSPset_seed_hsp (src/comparison/hspset.c:933-997):
// HSP_trim_ends (:850-878), HSP_init (:725-745), HSP_extend with masking forbidden and
// then ignored (:747-812), the threshold of HSP_store (:893-894) and HSP_find_cobs
// (:426-441) -- for every seed of one comparison in one launch, one thread per seed.
// Each seed is an independent, early-exit scan along one diagonal; the sequences are
// pre-encoded once per call to substitution-matrix indices (for a translated target:
// the matrix index of the codon starting at every position, and the OR of its three
// mask bytes), so a score is one shared-memory load.  HBM/L2-bound byte work: a seed
// reads 2 x (extension length) code bytes; nothing is written but 32 B per seed.
// The diagonal horizon is sequential over the seed list and stays with the caller.
#pragma once
#include "c4b_common.cuh"

namespace c4b {

// symbols -> matrix indices.  kind: 0 = 1:1 through `index`; 1 = translated codon
// starting at each position (positions n-2.. get 24).  *bad is raised by a symbol the
// matrix does not know (the reference would read out of bounds, submat.c:26-55).
__global__ void hsp_encode_kernel(const uint8_t *__restrict__ seq, int n, const uint8_t *__restrict__ index,
                                  const uint8_t *__restrict__ nt2d, const uint8_t *__restrict__ codon_aa,
                                  int translated, uint8_t *__restrict__ code, const uint8_t *__restrict__ mask,
                                  uint8_t *__restrict__ mask_out, int *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!translated) {
        const uint8_t c = index[seq[i]];
        if (c >= 24) atomicOr(bad, 1);
        code[i] = c >= 24 ? 0 : c;
        if (mask_out) mask_out[i] = mask ? mask[i] : 0;
    } else {
        uint8_t c = 24, m = 0;
        if (i + 2 < n) {
            const uint8_t aa = codon_aa[nt2d[seq[i]] | (nt2d[seq[i + 1]] << 4) | (nt2d[seq[i + 2]] << 8)];
            c = index[aa];
            if (c >= 24) { atomicOr(bad, 1); c = 0; }
            if (mask) m = mask[i] | mask[i + 1] | mask[i + 2];
        }
        code[i] = c;
        if (mask_out) mask_out[i] = m;
    }
}

struct HspArgs {
    const uint8_t *qc, *tc, *qm, *tm;  // codes and (already combined) masks; masks may be null
    int ql, tl, qadv, tadv;
    int seedlen, dropoff, threshold;
};

__global__ void hsp_extend_kernel(const HspArgs A, const int32_t *__restrict__ matrix, int n,
                                  const c4b_hsp_seed *__restrict__ seeds, c4b_hsp *__restrict__ out) {
    __shared__ int16_t sm[24 * 24];
    for (int k = threadIdx.x; k < 24 * 24; k += blockDim.x) sm[k] = (int16_t)matrix[k];
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int tadv = A.tadv, qadv = A.qadv;
    auto score_at = [&](int qp, int tp) -> int { return sm[A.qc[qp] * 24 + A.tc[tp]]; };
    auto masked = [&](int qp, int tp) -> bool { return (A.qm && A.qm[qp]) || (A.tm && A.tm[tp]); };

    c4b_hsp h;
    h.query_start = seeds[g].query_start;
    h.target_start = seeds[g].target_start;
    h.length = A.seedlen;
    h.score = 0; h.cobs = 0; h.stored = 0; h.status = 0;
    // HSP_trim_ends: drop non-positive ends of the seed
    {
        int i = 0;
        for (; i < h.length; ++i) {
            if (score_at(h.query_start, h.target_start) > 0) break;
            h.query_start += qadv;
            h.target_start += tadv;
        }
        h.length -= i;
        int qp = h.query_start + (h.length - 1) * qadv, tp = h.target_start + (h.length - 1) * tadv;
        while (h.length > 0) {
            if (score_at(qp, tp) > 0) break;
            --h.length;
            qp -= qadv;
            tp -= tadv;
        }
    }
    // HSP_init: seed score
    for (int i = 0, qp = h.query_start, tp = h.target_start; i < h.length; ++i, qp += qadv, tp += tadv)
        h.score += score_at(qp, tp);
    if (h.score < 0) {
        h.status = 1;  // "Initial HSP score less than zero" is fatal in the reference
        h.target_end = h.target_start + h.length * tadv;
        out[g] = h;
        return;
    }
    // HSP_extend: left then right; stop when the running score drops below zero or by
    // `dropoff` from its maximum; ties move the maximum outwards (maxscore <= score)
    auto extend = [&](bool forbid_masked) {
        int score = h.score, maxscore = h.score;
        int qp = h.query_start - qadv, tp = h.target_start - tadv, maxext = 0;
        for (int ext = 1; qp >= 0 && tp >= 0; ++ext) {
            if (forbid_masked && masked(qp, tp)) break;
            score += score_at(qp, tp);
            if (maxscore <= score) { maxscore = score; maxext = ext; }
            else if (score < 0 || maxscore - score >= A.dropoff) break;
            qp -= qadv;
            tp -= tadv;
        }
        qp = h.query_start + h.length * qadv;
        tp = h.target_start + h.length * tadv;
        h.query_start -= maxext * qadv;
        h.target_start -= maxext * tadv;
        h.length += maxext;
        score = maxscore;
        maxext = 0;
        for (int ext = 1; qp + qadv <= A.ql && tp + tadv <= A.tl; ++ext) {
            if (forbid_masked && masked(qp, tp)) break;
            score += score_at(qp, tp);
            if (maxscore <= score) { maxscore = score; maxext = ext; }
            else if (score < 0 || maxscore - score >= A.dropoff) break;
            qp += qadv;
            tp += tadv;
        }
        h.score = maxscore;
        h.length += maxext;
    };
    extend(true);
    if (h.score >= A.threshold) extend(false);
    h.target_end = h.target_start + h.length * tadv;
    h.stored = h.score >= A.threshold ? 1 : 0;
    if (h.stored) {  // HSP_find_cobs: first position where the prefix score reaches half
        int score = 0, i = 0;
        for (int qp = h.query_start, tp = h.target_start; i < h.length; ++i, qp += qadv, tp += tadv) {
            score += score_at(qp, tp);
            if (score >= (h.score >> 1)) break;
        }
        h.cobs = i;
    }
    out[g] = h;
}

}  // namespace c4b
