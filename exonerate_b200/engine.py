"""ctypes binding of libc4b200.so -- the reference-side stub a maintainer would
write, used by the tests and bench.py.

The names mirror the reference's Optimal API (src/c4/optimal.h:49-65):
Optimal.find_score / Optimal.find_path, batched.  There is NO fallback: if the
CUDA library is missing or no device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("C4B_LIB", os.path.join(_HERE, "libc4b200.so"))  # override: A/B kernel builds

EXPORTS = [
    "c4b_abi_version", "c4b_last_error", "c4b_engine_create", "c4b_engine_destroy",
    "c4b_engine_set_stream", "c4b_engine_forget_buffers", "c4b_engine_forget_buffer", "c4b_engine_kernel_launches", "c4b_find_score_batch",
    "c4b_find_path_batch", "c4b_batch_create", "c4b_batch_run", "c4b_batch_fetch",
    "c4b_batch_cells", "c4b_batch_ops_needed", "c4b_batch_device_results", "c4b_batch_last_fill_ms", "c4b_batch_kernel_name", "c4b_batch_description", "c4b_batch_destroy",
    "c4b_group_create", "c4b_group_destroy", "c4b_group_size", "c4b_group_find_score_batch",
    "c4b_group_find_path_batch", "c4b_group_kernel_launches", "c4b_free",
    "c4b_viterbi_calculate", "c4b_viterbi_calculate_cells", "c4b_hsp_extend_batch", "c4b_model_specialise", "c4b_span_integrate", "c4b_span_score_batch",
]

_lib = None


class C4BError(RuntimeError):
    pass


def load_library():
    """dlopen the in-tree CUDA library; fails loudly when it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise C4BError("libc4b200.so is not built (run __graft_entry__.build()); "
                       "there is no CPU fallback for the C4 fill")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.c4b_abi_version.restype = C.c_int
    lib.c4b_last_error.restype = C.c_char_p
    lib.c4b_engine_create.argtypes = [C.c_int, P(C.c_void_p)]
    lib.c4b_engine_destroy.argtypes = [C.c_void_p]
    lib.c4b_engine_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.c4b_engine_kernel_launches.argtypes = [C.c_void_p]
    lib.c4b_engine_kernel_launches.restype = C.c_int64
    lib.c4b_find_score_batch.argtypes = [C.c_void_p, P(abi.Model), P(abi.Scoring), C.c_int32,
                                         P(abi.Pair), C.c_void_p]
    lib.c4b_find_path_batch.argtypes = [C.c_void_p, P(abi.Model), P(abi.Scoring), C.c_int32, P(abi.Pair),
                                        C.c_int32, P(abi.Result), C.c_void_p, C.c_int64]
    lib.c4b_batch_create.argtypes = [C.c_void_p, P(abi.Model), P(abi.Scoring), C.c_int32, P(abi.Pair),
                                     C.c_int, P(C.c_void_p)]
    lib.c4b_batch_run.argtypes = [C.c_void_p, C.c_int32]
    lib.c4b_batch_fetch.argtypes = [C.c_void_p, P(abi.Result), C.c_void_p, C.c_int64]
    lib.c4b_batch_device_results.argtypes = [C.c_void_p]
    lib.c4b_batch_device_results.restype = C.c_void_p
    lib.c4b_batch_ops_needed.argtypes = [C.c_void_p]
    lib.c4b_batch_ops_needed.restype = C.c_int64
    lib.c4b_batch_cells.argtypes = [C.c_void_p]
    lib.c4b_batch_cells.restype = C.c_int64
    lib.c4b_viterbi_calculate_cells.argtypes = [C.c_void_p, P(abi.Model), P(abi.Scoring), P(abi.Pair), C.c_int,
                                                C.c_void_p, C.c_void_p, P(abi.Result), C.c_void_p, C.c_int64]
    lib.c4b_viterbi_calculate_cells.restype = C.c_int
    lib.c4b_model_specialise.argtypes = [P(abi.Model), C.c_int32, C.c_int32, P(C.c_int64)]
    lib.c4b_model_specialise.restype = C.c_int
    lib.c4b_span_integrate.argtypes = [C.c_void_p] + [P(C.c_int32)] * 5
    lib.c4b_span_integrate.restype = C.c_int
    lib.c4b_engine_forget_buffers.argtypes = [C.c_void_p]
    lib.c4b_engine_forget_buffers.restype = None
    lib.c4b_hsp_extend_batch.argtypes = [C.c_void_p, P(abi.Scoring), P(abi.HspParam), C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                         C.c_void_p]
    lib.c4b_hsp_extend_batch.restype = C.c_int
    lib.c4b_batch_last_fill_ms.argtypes = [C.c_void_p]
    lib.c4b_batch_last_fill_ms.restype = C.c_double
    lib.c4b_batch_kernel_name.argtypes = [C.c_void_p]
    lib.c4b_batch_kernel_name.restype = C.c_char_p
    lib.c4b_group_create.argtypes = [C.c_int, P(C.c_int), P(C.c_void_p)]
    lib.c4b_group_destroy.argtypes = [C.c_void_p]
    lib.c4b_group_size.argtypes = [C.c_void_p]
    lib.c4b_group_kernel_launches.argtypes = [C.c_void_p]
    lib.c4b_group_kernel_launches.restype = C.c_int64
    lib.c4b_group_find_score_batch.argtypes = [C.c_void_p, P(abi.Model), P(abi.Scoring), C.c_int32, P(abi.Pair),
                                               C.c_void_p]
    lib.c4b_group_find_path_batch.argtypes = [C.c_void_p, P(abi.Model), P(abi.Scoring), C.c_int32, P(abi.Pair),
                                              C.c_int32, P(abi.Result), P(C.c_void_p), P(C.c_int64)]
    lib.c4b_free.argtypes = [C.c_void_p]
    lib.c4b_batch_description.argtypes = [C.c_void_p]
    lib.c4b_batch_description.restype = C.c_char_p
    lib.c4b_batch_destroy.argtypes = [C.c_void_p]
    lib.c4b_viterbi_calculate.argtypes = [C.c_void_p, P(abi.Model), P(abi.Scoring), P(abi.Pair), C.c_int,
                                          P(abi.Result), C.c_void_p, C.c_int64]
    if lib.c4b_abi_version() != abi.ABI_VERSION:
        raise C4BError("libc4b200.so ABI version mismatch")
    _lib = lib
    return lib


def _check(lib, rc, what):
    if rc != 0:
        raise C4BError("%s failed (rc=%d): %s" % (what, rc, lib.c4b_last_error().decode()))


class PairSet:
    """n query x target lattices as a contiguous c4b_pair array (host buffers)."""

    def __init__(self, queries, targets, splice=None, blocked=None, regions=None, pinned=False):
        """pinned=True: the caller guarantees that every sequence buffer (numpy arrays, used in place)
        lives in page-locked memory -- e.g. rows of torch.empty(..., pin_memory=True).numpy() -- and
        stays valid until the batch has been fetched (C4B_PAIR_BUFFERS_PINNED)."""
        assert len(queries) == len(targets)
        n = len(queries)
        self.n = n
        self._keep = []
        self.array = (abi.Pair * max(n, 1))()
        cache = {}

        def buf(s):
            key = id(s)
            if key not in cache:
                if isinstance(s, np.ndarray):  # used in place (must be contiguous uint8)
                    assert s.dtype == np.uint8 and s.flags["C_CONTIGUOUS"]
                    cache[key] = s
                else:
                    raw = s.encode() if isinstance(s, str) else bytes(s)
                    cache[key] = np.frombuffer(raw, dtype=np.uint8).copy()
                self._keep.append(cache[key])
            return cache[key]

        for k in range(n):
            q, t = buf(queries[k]), buf(targets[k])
            p = self.array[k]
            p.query, p.target = q.ctypes.data, t.ctypes.data
            p.query_len, p.target_len = len(q), len(t)
            reg = regions[k] if regions else (0, 0, len(q), len(t))
            p.query_start, p.target_start, p.query_length, p.target_length = reg
            if pinned:
                assert isinstance(queries[k], np.ndarray) and isinstance(targets[k], np.ndarray)
                p.reserved |= abi.PAIR_BUFFERS_PINNED
            if splice and splice[k] is not None:
                arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in splice[k]]
                self._keep.append(arrs)
                for i, a in enumerate(arrs):
                    p.splice[i] = a.ctypes.data
            if blocked and blocked[k]:
                pts = sorted(set((tj, qi) for qi, tj in blocked[k]))
                bq = np.array([qi for tj, qi in pts], dtype=np.int32)
                bt = np.array([tj for tj, qi in pts], dtype=np.int32)
                self._keep.append((bq, bt))
                p.blocked_query_pos, p.blocked_target_pos, p.n_blocked = bq.ctypes.data, bt.ctypes.data, len(pts)
        self.cells = sum(int(self.array[k].query_length) * int(self.array[k].target_length) for k in range(n))
        # bytes the engine copies to the device: every distinct sequence once, plus one
        # packed 4 x int8 splice word per target position where splice arrays are given
        self.h2d_bytes = sum(a.nbytes for a in cache.values())
        if splice:
            self.h2d_bytes += 4 * sum(int(self.array[k].target_length) for k in range(n) if splice[k] is not None)


def results_to_list(results, ops, n):
    out = []
    for k in range(n):
        r = results[k]
        o = int(r.ops_offset)
        out.append({"score": r.score,
                    "region": [r.query_start, r.target_start, r.query_end - r.query_start,
                               r.target_end - r.target_start],
                    "ops": [(int(ops[2 * (o + i)]), int(ops[2 * (o + i) + 1])) for i in range(r.n_ops)],
                    "status": r.status})
    return out


class Batch:
    """A staged (HBM-resident) batch: create once, run many times, fetch."""

    def __init__(self, engine, model, scoring, pairs, want_path=True):
        self.engine, self.lib = engine, engine.lib
        self.pairs, self.want_path = pairs, want_path
        self.h = C.c_void_p()
        _check(self.lib, self.lib.c4b_batch_create(engine.h, C.byref(model), C.byref(scoring), pairs.n,
                                                   pairs.array, int(want_path), C.byref(self.h)),
               "c4b_batch_create")

    def run(self, threshold=abi.IMPOSSIBLY_LOW_SCORE):
        _check(self.lib, self.lib.c4b_batch_run(self.h, threshold), "c4b_batch_run")

    def fetch(self, ops_capacity=None):
        n = self.pairs.n
        results = (abi.Result * max(n, 1))()
        if ops_capacity is None:
            ops_capacity = sum(int(self.pairs.array[k].query_length) + int(self.pairs.array[k].target_length) + 4
                               for k in range(n)) if self.want_path else 0
        ops = np.zeros(2 * max(ops_capacity, 1), dtype=np.int32)
        _check(self.lib, self.lib.c4b_batch_fetch(self.h, results, ops.ctypes.data, ops_capacity),
               "c4b_batch_fetch")
        return results, ops

    @property
    def cells(self):
        return self.lib.c4b_batch_cells(self.h)

    def ops_needed(self):
        n = self.lib.c4b_batch_ops_needed(self.h)
        if n < 0:
            raise C4BError("c4b_batch_ops_needed: " + self.lib.c4b_last_error().decode())
        return n

    @property
    def kernel_name(self):
        return self.lib.c4b_batch_kernel_name(self.h).decode()

    @property
    def description(self):
        return self.lib.c4b_batch_description(self.h).decode()

    def last_fill_ms(self):
        return self.lib.c4b_batch_last_fill_ms(self.h)

    def device_results(self):
        """The c4b_result[n] array in HBM as an object exposing
        __cuda_array_interface__ (int32 [n, 10]); e.g. torch.as_tensor(x, device="cuda")."""
        ptr = self.lib.c4b_batch_device_results(self.h)
        if not ptr:
            raise C4BError("no device results (run the batch first; score-only batches have none)")

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": (self.pairs.n, 10), "typestr": "<i4", "data": (ptr, False),
                                      "version": 3, "strides": None}
        v._owner = self
        return v

    def close(self):
        if self.h:
            self.lib.c4b_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        _check(self.lib, self.lib.c4b_engine_create(device, C.byref(self.h)), "c4b_engine_create")

    def kernel_launches(self):
        return self.lib.c4b_engine_kernel_launches(self.h)

    def close(self):
        if self.h:
            self.lib.c4b_engine_destroy(self.h)
            self.h = C.c_void_p()


class Group:
    """c4b_group: one batch sharded over several GPUs of one box inside the C library (one engine and
    one host thread per device, lattices dealt by cost); same results as one Engine."""

    def __init__(self, devices=None):
        self.lib = load_library()
        self.h = C.c_void_p()
        devs = list(devices) if devices else []
        arr = (C.c_int * max(1, len(devs)))(*devs)
        _check(self.lib, self.lib.c4b_group_create(len(devs), arr, C.byref(self.h)), "c4b_group_create")

    @property
    def size(self):
        return self.lib.c4b_group_size(self.h)

    def kernel_launches(self):
        return self.lib.c4b_group_kernel_launches(self.h)

    def find_score(self, model, scoring, pairs):
        scores = np.zeros(max(pairs.n, 1), dtype=np.int32)
        _check(self.lib, self.lib.c4b_group_find_score_batch(self.h, C.byref(model), C.byref(scoring), pairs.n,
                                                             pairs.array, scores.ctypes.data),
               "c4b_group_find_score_batch")
        return [int(x) for x in scores[:pairs.n]]

    def find_path_raw(self, model, scoring, pairs, threshold=abi.IMPOSSIBLY_LOW_SCORE):
        """(c4b_result array, int32 ops array copied out of the library's buffer)"""
        results = (abi.Result * max(pairs.n, 1))()
        ops_p, n_ops = C.c_void_p(), C.c_int64()
        _check(self.lib, self.lib.c4b_group_find_path_batch(self.h, C.byref(model), C.byref(scoring), pairs.n,
                                                            pairs.array, threshold, results, C.byref(ops_p),
                                                            C.byref(n_ops)), "c4b_group_find_path_batch")
        ops = np.ctypeslib.as_array(C.cast(ops_p, C.POINTER(C.c_int32)), shape=(2 * max(n_ops.value, 1),)).copy()
        self.lib.c4b_free(ops_p)
        return results, ops

    def find_path(self, model, scoring, pairs, threshold=abi.IMPOSSIBLY_LOW_SCORE):
        results, ops = self.find_path_raw(model, scoring, pairs, threshold)
        return results_to_list(results, ops, pairs.n)

    def close(self):
        if self.h:
            self.lib.c4b_group_destroy(self.h)
            self.h = C.c_void_p()


class Optimal:
    """Optimal_create / Optimal_find_score / Optimal_find_path over batches
    (src/c4/optimal.h:49-65)."""

    def __init__(self, engine, model, scoring):
        self.engine, self.model, self.scoring = engine, model, scoring

    def find_score(self, pairs):
        lib = self.engine.lib
        scores = np.zeros(max(pairs.n, 1), dtype=np.int32)
        _check(lib, lib.c4b_find_score_batch(self.engine.h, C.byref(self.model), C.byref(self.scoring),
                                             pairs.n, pairs.array, scores.ctypes.data), "c4b_find_score_batch")
        return [int(x) for x in scores[:pairs.n]]

    def find_path_raw(self, pairs, threshold=abi.IMPOSSIBLY_LOW_SCORE, ops_capacity=None, out=None):
        """c4b_find_path_batch as is: (c4b_result array, int32 ops array).
        `out` = (results, ops) buffers to reuse between calls."""
        lib = self.engine.lib
        n = pairs.n
        if out is not None:
            results, ops = out
            ops_capacity = len(ops) // 2
        else:
            results = (abi.Result * max(n, 1))()
            if ops_capacity is None:
                ops_capacity = sum(int(pairs.array[k].query_length) + int(pairs.array[k].target_length) + 4
                                   for k in range(n))
            ops = np.empty(2 * max(ops_capacity, 1), dtype=np.int32)
        _check(lib, lib.c4b_find_path_batch(self.engine.h, C.byref(self.model), C.byref(self.scoring), n,
                                            pairs.array, threshold, results, ops.ctypes.data, ops_capacity),
               "c4b_find_path_batch")
        return results, ops

    def find_path(self, pairs, threshold=abi.IMPOSSIBLY_LOW_SCORE, ops_capacity=None):
        results, ops = self.find_path_raw(pairs, threshold, ops_capacity)
        return results_to_list(results, ops, pairs.n)


def viterbi_calculate_cells(engine, model, scoring, pairs, mode, start_cells=None, end_cells=None, max_ops=1 << 14):
    """c4b_viterbi_calculate_cells on lattice 0 of `pairs` (BSDP derived models: START-cell
    table in, END-cell table out; int32 numpy arrays).  Returns results_to_list()[0]."""
    lib = engine.lib
    res = (abi.Result * 1)()
    ops = np.zeros(2 * max_ops, dtype=np.int32)
    _check(lib, lib.c4b_viterbi_calculate_cells(engine.h, C.byref(model), C.byref(scoring), pairs.array, mode,
                                                start_cells.ctypes.data if start_cells is not None else None,
                                                end_cells.ctypes.data if end_cells is not None else None,
                                                res, ops.ctypes.data, max_ops), "c4b_viterbi_calculate_cells")
    return results_to_list(res, ops, 1)[0]


class HSPset:
    """HSPset_create / HSPset_seed_hsp / HSPset_finalise (src/comparison/hspset.h:205-229)
    for one query x target comparison.  Seeds are collected; finalise() extends all of
    them on the device in one c4b_hsp_extend_batch call and then replays the diagonal
    horizon of HSPset_seed_hsp (hspset.c:933-972,991-996; seed_repeat 1, no hspfilter)
    over the results in seed order, which reproduces hsp_list exactly."""

    def __init__(self, engine, scoring, param, query, target, query_mask=None, target_mask=None):
        self.engine, self.scoring, self.param = engine, scoring, param
        self.q = np.frombuffer(query.encode() if isinstance(query, str) else bytes(query), dtype=np.uint8).copy()
        self.t = np.frombuffer(target.encode() if isinstance(target, str) else bytes(target), dtype=np.uint8).copy()
        self.qm = None if query_mask is None else np.ascontiguousarray(query_mask, dtype=np.uint8)
        self.tm = None if target_mask is None else np.ascontiguousarray(target_mask, dtype=np.uint8)
        self.seeds = []
        self.hsp_list = None

    def seed_hsp(self, query_start, target_start):
        assert self.hsp_list is None, "HSPset already finalised"
        self.seeds.append((int(query_start), int(target_start)))

    def extend_all(self):
        """per-seed device results (abi.Hsp array), before the horizon"""
        lib = self.engine.lib
        n = len(self.seeds)
        sd = (abi.HspSeed * max(1, n))(*[abi.HspSeed(a, b) for a, b in self.seeds])
        out = (abi.Hsp * max(1, n))()
        _check(lib, lib.c4b_hsp_extend_batch(self.engine.h, C.byref(self.scoring), C.byref(self.param),
                                             self.q.ctypes.data, len(self.q),
                                             self.qm.ctypes.data if self.qm is not None else None,
                                             self.t.ctypes.data, len(self.t),
                                             self.tm.ctypes.data if self.tm is not None else None,
                                             n, sd, out), "c4b_hsp_extend_batch")
        return out

    def finalise(self):
        ext = self.extend_all()
        tadv = 3 if self.param.match_kind in (abi.CALC_MATCH_1_3, abi.CALC_MATCH_3_3) else 1
        qadv = 3 if self.param.match_kind in (abi.CALC_MATCH_3_1, abi.CALC_MATCH_3_3) else 1
        ql = len(self.q)
        horizon = {}
        self.hsp_list = []
        for k, (qs, ts) in enumerate(self.seeds):
            key = ((ts * qadv - qs * tadv + ql) % ql, qs % qadv, ts % tadv)
            if ts < horizon.get(key, 0):   # skipped before HSP_init ever sees it (hspset.c:960-972)
                continue
            if ext[k].status != 0:   # the reference aborts here (HSP_init, hspset.c:740-743)
                raise C4BError("Initial HSP score less than zero for seed (%d, %d)" % (qs, ts))
            horizon[key] = ext[k].target_end
            if ext[k].stored:
                h = ext[k]
                self.hsp_list.append([h.query_start, h.target_start, h.length, h.score, h.cobs])
        return self.hsp_list


def span_score_batch(engine, src_model, dst_model, scoring, pairs, src_regions, dst_regions, spans):
    """c4b_span_score_batch: SAR_Span_find_score (src/bsdp/sar.c:898-917) for a list of span edges.
    pairs: a PairSet whose lattice k carries the sequences / splice arrays of edge k; src_regions /
    dst_regions: (query_start, target_start, query_length, target_length); spans: (min_q, max_q, min_t, max_t)."""
    lib = engine.lib
    n = pairs.n
    jobs = (abi.SpanJob * max(n, 1))()
    for k in range(n):
        for side, reg in ((jobs[k].src, src_regions[k]), (jobs[k].dst, dst_regions[k])):
            C.memmove(C.byref(side), C.byref(pairs.array[k]), C.sizeof(abi.Pair))
            side.query_start, side.target_start, side.query_length, side.target_length = reg
        for l in range(4):
            jobs[k].span[l] = spans[k][l]
    scores = np.zeros(max(n, 1), dtype=np.int32)
    lib.c4b_span_score_batch.argtypes = [C.c_void_p, C.POINTER(abi.Model), C.POINTER(abi.Model), C.POINTER(abi.Scoring),
                                         C.c_int32, C.POINTER(abi.SpanJob), C.c_void_p]
    _check(lib, lib.c4b_span_score_batch(engine.h, C.byref(src_model), C.byref(dst_model), C.byref(scoring), n, jobs,
                                         scores.ctypes.data), "c4b_span_score_batch")
    return [int(x) for x in scores[:n]]


def span_integrate(engine, src_scores, src_region, dst_region, span):
    """c4b_span_integrate: Heuristic_Span_integrate's scan (src/bsdp/heuristic.c:589-678) on the
    device.  src_scores: (src_ql+1)*(src_tl+1) ints; regions {q_start,t_start,q_len,t_len};
    span {min_q,max_q,min_t,max_t}.  Returns the flat list of (query_pos, target_pos) per dst cell."""
    import numpy as np
    lib = engine.lib
    I = C.POINTER(C.c_int32)
    sc = np.ascontiguousarray(src_scores, dtype=np.int32)
    src = np.asarray(src_region, dtype=np.int32)
    dst = np.asarray(dst_region, dtype=np.int32)
    sp = np.asarray(span, dtype=np.int32)
    out = np.full(2 * (int(dst[2]) + 1) * (int(dst[3]) + 1), -7, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(I)
    _check(lib, lib.c4b_span_integrate(engine.h, p(sc), p(src), p(dst), p(sp), p(out)), "c4b_span_integrate")
    return out.tolist()

