// generic_types.h -- lattice descriptors of the table-driven path, shared by the
// nvcc-built interpreter kernel (generic_wavefront.cuh) and the NVRTC-built
// model-specialised kernel (generic_jit_kernel.cuh).  Plain structs only: this
// file is also handed to NVRTC as an in-memory header.
#pragma once
#include <stdint.h>

#include "c4b200.h"

namespace c4b {

constexpr int kMaxCell = 1 + C4B_MAX_SHADOW_SLOTS + 2;

struct GenPair {
    const uint8_t *q, *t;  // raw symbol bytes, whole sequences
    const int32_t *splice[4];
    const int32_t *blk_q, *blk_t;
    int32_t n_blocked;
    int32_t q_start, t_start, Q, T;  // region origin + extents
    int32_t blk_dq, blk_dt;          // blocked coordinates are relative to (q_start-blk_dq, ...)
    uint8_t *tb;                     // PATH: (Q+1)*(T+1)*S winning transition ids (0xFF = none)
    // cell callbacks of BSDP derived models, as tables ((Q+1)*(T+1) cells of 1 + n_shadow_slots ints):
    const int32_t *start_cells;      // what cell_start_func returns per cell (viterbi.c:727-741), or null
    int32_t *end_cells;              // END's cell wherever END is reached (cell_end_func's input), or null
    int64_t out_index;
    // PATH record layout: 0 = by anti-diagonal, one byte per state (GEN_TB_CELL); r > 0 = the systolic
    // kernel's skewed bit-packed layout with r rows per lane (GEN_TBS_CHUNK)
    int32_t tb_rows, tb_chunk;
    // SubOpt blocked cells for the systolic kernel (JIT_SYS_BLK): per lane strip (strip * 32 + lane) the
    // blocked MATCH destination cells of THIS lattice as {column, row mask} sorted by column; blk_off has
    // one entry per lane strip + 1.  Null when the lattice has none (or on the other kernels, which
    // search blk_q / blk_t instead).
    const int2 *blk;
    const int32_t *blk_off;
};

struct GenOut {
    int32_t score, end_i, end_j, start_i, start_j, flags;
};

// Column windows of the systolic PATH pass (generic_jit_systolic.cuh, JIT_SYS_WIN), one per lattice
// of the launch: pass 1 leaves a checkpoint of the register lattice after every `wcols` columns, the
// PATH pass refills columns [c0, c1] of strips 0 .. nsweeps - 1 from the checkpoint left of c0.
struct GenWin {
    int32_t *ck;         // [window boundary][strip][word][lane]
    int32_t wcols;       // window width, a power of two
    int32_t c0, c1;      // columns of this refill (c0 a multiple of wcols; c1 = the cursor's column)
    int32_t nsweeps;     // strips down to the cursor's (0: nothing to do for this lattice)
    int32_t reserved;
};

// traceback cursor of a windowed lattice between rounds (generic_window_walk_kernel)
struct GenWalk {
    int32_t i, j, state;     // next: look up the winner of `state` at cell (i, j)
    int32_t n_runs, last_t, status, done, reserved;
};

struct GenTables {
    c4b_model model;
    c4b_scoring scoring;
};

enum { GEN_SCORE = 0, GEN_PATH = 1, GEN_REGION = 2 };

// PATH traceback bytes (winning transition id per state per cell, viterbi.c:220-227 keeps
// pointers) are laid out by ANTI-DIAGONAL: cell (i,j) at [(i+j)*(Q+1) + i][S].  The fill walks
// diagonals with one thread per row, so a warp's stores land in consecutive S-byte groups
// (row-major cells would put every thread in its own 32-byte sector: 32x the write traffic).
#define GEN_TB_CELL(i, j, Q, S) (((size_t)((i) + (j)) * (size_t)((Q) + 1) + (size_t)(i)) * (size_t)(S))
#define GEN_TB_BYTES(Q, T, S) ((size_t)((Q) + (T) + 1) * (size_t)((Q) + 1) * (size_t)(S))
// the systolic layout (generic_jit_systolic.cuh): [strip][step][lane][CH bytes], CH bytes = the R
// rows of a lane at one column, row r at bit r * row_bits, state s at its bit offset inside the row:
// the 1-based rank of the winning transition among those entering s (0 = unset).  Chunk of cell
// (i, j) in a lattice with T target positions:
#define GEN_TBS_CHUNK(i, j, T, R, CH)                                                                    \
    ((((size_t)((i) / (32 * (R))) * (size_t)((T) + 32) + (size_t)((j) + ((i) / (R)) % 32)) * 32 +        \
      (size_t)(((i) / (R)) % 32)) * (size_t)(CH))
#define GEN_TBS_BYTES(Q, T, R, CH) \
    ((size_t)(((Q) + 32 * (R)) / (32 * (R))) * (size_t)((T) + 32) * 32 * (size_t)(CH))

}  // namespace c4b
