// c4b_common.cuh -- shared device/host definitions of libc4b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "c4b200.h"

namespace c4b {

// "unset" / invalid-candidate sentinels.  C4_IMPOSSIBLY_LOW_SCORE is the
// reference's reset value (src/c4/c4.h:29); NEG2 is our own "transition not
// valid here" input, low enough never to win against a reachable score and
// high enough that NEG2 + penalties cannot wrap int32.
constexpr int32_t LOW = C4B_IMPOSSIBLY_LOW_SCORE;
constexpr int32_t NEG2 = -1900000000;
constexpr int kTargetNone = 24;  // "no symbol" column code (lattice column 0)
constexpr int kPadClass = 7;     // PRMT query class of rows that hold no query symbol

void set_error(const std::string &msg);

#define C4B_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t err__ = (call);                                                 \
        if (err__ != cudaSuccess) {                                                 \
            c4b::set_error(std::string(#call) + ": " + cudaGetErrorString(err__)); \
            return -1;                                                              \
        }                                                                           \
    } while (0)

// ---- affine systolic path --------------------------------------------------
// One query x target lattice as the fill kernel sees it.  rows = Q+1 lattice
// rows (row 0 = "no query symbol consumed"), cols = T+1.
struct AffPair {
    const uint8_t *q;  // per query position: PRMT class (0..6) or matrix row (0..23)
    const uint8_t *t;  // per target position: matrix column code 0..23
    int32_t Q, T;
    uint32_t *tb;      // traceback nibbles, skewed layout [sweep][step][lane][R/8]; may be null
    int2 *top0, *top1; // sweep hand-off rows {M,I}[T+1], ping-pong; null when one sweep
    int64_t out_index; // which result slot this lattice reports into
    // SubOpt blocked cells (int32 fill only): entries {column, row mask} of lane strip k are
    // blk[blk_off[k] .. blk_off[k+1]), sorted by column; blk_off null = nothing blocked;
    // blk_j0 = lattice column 0 in the list's coordinates (non-zero for a banded refill)
    const int2 *blk;
    const int32_t *blk_off;
    int32_t blk_j0, blk_pad;
};

struct AffOut {
    int32_t best, end_i, end_j, flags;
};

struct AffModel {
    int32_t openD, extD, openI, extI; // penalties on the delete / insert chains
    int32_t start_scope, end_scope;
    int32_t score_mode;               // 0 = PRMT classes, 1 = smem matrix rows
    int32_t one;                      // run-time 1 (see add_open)
    // transition ids of the closed model, in template order
    int32_t tDD, tII, tMD, tMI, tMM, tSM, tDM, tIM, tME;
};

}  // namespace c4b
