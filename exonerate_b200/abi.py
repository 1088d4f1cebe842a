"""ctypes mirror of include/c4b200.h (struct layouts + constants only).

No computation lives here.  The product library (libc4b200.so) and the test
oracle (oracle/_ref/liboracle.so) both speak these structs.
"""
import ctypes as C

ABI_VERSION = 1
IMPOSSIBLY_LOW_SCORE = -987654321
MAX_STATES = 32
MAX_TRANSITIONS = 64
MAX_CALCS = 32
MAX_SHADOW_SLOTS = 4
SUBMAT_N = 24

SCOPE_ANYWHERE, SCOPE_EDGE, SCOPE_QUERY, SCOPE_TARGET, SCOPE_CORNER = range(5)
(LABEL_NONE, LABEL_MATCH, LABEL_GAP, LABEL_NER, LABEL_5SS, LABEL_3SS, LABEL_INTRON,
 LABEL_SPLIT_CODON, LABEL_FRAMESHIFT) = range(9)
PROTECT_NONE, PROTECT_OVERFLOW, PROTECT_UNDERFLOW = 0, 1, 2
(CALC_CONST, CALC_MATCH_DNA, CALC_MATCH_PROTEIN, CALC_MATCH_1_3, CALC_MATCH_3_1, CALC_MATCH_3_3,
 CALC_SPLICE_PRE, CALC_SPLICE_POST, CALC_PHASE1_POST, CALC_PHASE2_POST) = range(10)
SPLICE_5_FORWARD, SPLICE_3_FORWARD, SPLICE_5_REVERSE, SPLICE_3_REVERSE = range(4)
MODE_FIND_SCORE, MODE_FIND_PATH, MODE_FIND_REGION = 0, 1, 2


class Calc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("protect", C.c_int32), ("param", C.c_int32 * 4)]


class Transition(C.Structure):
    _fields_ = [("input", C.c_int32), ("output", C.c_int32), ("advance_query", C.c_int32),
                ("advance_target", C.c_int32), ("calc", C.c_int32), ("label", C.c_int32)]


class Model(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("n_transitions", C.c_int32), ("n_calcs", C.c_int32),
                ("n_shadow_slots", C.c_int32), ("start_state", C.c_int32), ("end_state", C.c_int32),
                ("start_scope", C.c_int32), ("end_scope", C.c_int32),
                ("max_query_advance", C.c_int32), ("max_target_advance", C.c_int32),
                ("shadow_start", (C.c_uint8 * MAX_SHADOW_SLOTS) * MAX_STATES),
                ("transitions", Transition * MAX_TRANSITIONS),
                ("calcs", Calc * MAX_CALCS)]


class Scoring(C.Structure):
    _fields_ = [("dna_matrix", C.c_int32 * (SUBMAT_N * SUBMAT_N)),
                ("protein_matrix", C.c_int32 * (SUBMAT_N * SUBMAT_N)),
                ("dna_index", C.c_uint8 * 256), ("protein_index", C.c_uint8 * 256),
                ("nt2d", C.c_uint8 * 256), ("codon_aa", C.c_uint8 * 4096),
                ("min_intron", C.c_int32), ("max_intron", C.c_int32)]


PAIR_BUFFERS_STABLE, PAIR_BUFFERS_PINNED = 1, 2


class Pair(C.Structure):
    _fields_ = [("query", C.c_void_p), ("target", C.c_void_p),
                ("query_len", C.c_int32), ("target_len", C.c_int32),
                ("query_start", C.c_int32), ("target_start", C.c_int32),
                ("query_length", C.c_int32), ("target_length", C.c_int32),
                ("splice", C.c_void_p * 4),
                ("blocked_query_pos", C.c_void_p), ("blocked_target_pos", C.c_void_p),
                ("n_blocked", C.c_int32), ("reserved", C.c_int32)]


class SpanJob(C.Structure):   # c4b_span_job
    _fields_ = [("src", Pair), ("dst", Pair), ("span", C.c_int32 * 4)]


class Result(C.Structure):
    _fields_ = [("score", C.c_int32), ("query_start", C.c_int32), ("target_start", C.c_int32),
                ("query_end", C.c_int32), ("target_end", C.c_int32), ("n_ops", C.c_int32),
                ("ops_offset", C.c_int64), ("status", C.c_int32), ("reserved", C.c_int32)]


# ---- HSP seeding / extension (c4b_hsp_extend_batch) -------------------------
class HspParam(C.Structure):
    _fields_ = [("match_kind", C.c_int32), ("seedlen", C.c_int32), ("dropoff", C.c_int32),
                ("threshold", C.c_int32)]


class HspSeed(C.Structure):
    _fields_ = [("query_start", C.c_int32), ("target_start", C.c_int32)]


class Hsp(C.Structure):
    _fields_ = [("query_start", C.c_int32), ("target_start", C.c_int32), ("length", C.c_int32),
                ("score", C.c_int32), ("cobs", C.c_int32), ("stored", C.c_int32),
                ("target_end", C.c_int32), ("status", C.c_int32)]
