"""ctypes wrapper around oracle/_ref/libc4ref.so (TEST INFRASTRUCTURE ONLY).

libc4ref.so = the UNMODIFIED reference compiled in place by oracle/Makefile plus
oracle/ref_driver.c.  It exists only where /root/reference was present at build
time (this container); the built .so travels to the GPU box.  Used by
tests/golden/make_golden.py (golden vectors) and bench.py's reference arm.
Never imported by the product package.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libc4ref.so")

_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)


def available():
    return os.path.exists(LIB_PATH)


class RefPair:
    def __init__(self, lib, model, qid, qseq, tid, tseq):
        self.lib = lib
        self.h = lib.c4ref_pair_open(model.h, qid.encode(), qseq.encode(), tid.encode(), tseq.encode())

    def score(self, use_subopt=False):
        return self.lib.c4ref_pair_score(self.h, int(use_subopt))

    def path(self, threshold=-987654321, use_subopt=False, add_to_subopt=False, max_ops=1 << 16):
        score = C.c_int()
        region = (C.c_int * 4)()
        ops = (C.c_int * (2 * max_ops))()
        n_ops = C.c_int()
        vulgar = C.c_char_p()
        cigar = C.c_char_p()
        vp, cp = C.c_void_p(), C.c_void_p()
        ok = self.lib.c4ref_pair_path(self.h, threshold, int(use_subopt), int(add_to_subopt),
                                      C.byref(score), region, ops, max_ops, C.byref(n_ops),
                                      C.byref(vp), C.byref(cp))
        if not ok:
            return None
        vulgar = C.string_at(vp.value).decode()
        cigar = C.string_at(cp.value).decode()
        self.lib.c4ref_free(vp)
        self.lib.c4ref_free(cp)
        n = n_ops.value
        assert n <= max_ops
        return {
            "score": score.value,
            "region": list(region),
            "ops": [(ops[2 * i], ops[2 * i + 1]) for i in range(n)],
            "vulgar": vulgar.rstrip("\n"),
            "cigar": cigar.rstrip("\n"),
        }

    def time_path(self, n_rep=1):
        score = C.c_int()
        t = self.lib.c4ref_pair_time_path(self.h, n_rep, C.byref(score))
        return t, score.value

    def close(self):
        self.lib.c4ref_pair_close(self.h)
        self.h = None


class RefModel:
    def __init__(self, lib, name, query_is_protein=False, target_is_protein=False, compiled=True):
        self.lib = lib
        self.name = name
        self.h = lib.c4ref_model_open(name.encode(), int(query_is_protein), int(target_is_protein), int(compiled))

    def dump(self):
        p = self.lib.c4ref_model_dump(self.h)
        s = C.string_at(p).decode()
        self.lib.c4ref_free(p)
        return s

    def pair(self, qseq, tseq, qid="query", tid="target"):
        return RefPair(self.lib, self, qid, qseq, tid, tseq)

    def close(self):
        self.lib.c4ref_model_close(self.h)
        self.h = None


def _load():
    lib = C.CDLL(LIB_PATH)
    lib.c4ref_session.argtypes = [C.c_int, C.POINTER(C.c_char_p), _CB, C.c_void_p]
    lib.c4ref_session.restype = C.c_int
    lib.c4ref_model_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.c4ref_model_open.restype = C.c_void_p
    lib.c4ref_model_close.argtypes = [C.c_void_p]
    lib.c4ref_model_dump.argtypes = [C.c_void_p]
    lib.c4ref_model_dump.restype = C.c_void_p
    lib.c4ref_free.argtypes = [C.c_void_p]
    lib.c4ref_pair_open.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
    lib.c4ref_pair_open.restype = C.c_void_p
    lib.c4ref_pair_close.argtypes = [C.c_void_p]
    lib.c4ref_pair_score.argtypes = [C.c_void_p, C.c_int]
    lib.c4ref_pair_score.restype = C.c_int
    lib.c4ref_pair_path.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.c4ref_pair_path.restype = C.c_int
    lib.c4ref_pair_time_path.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.c4ref_pair_time_path.restype = C.c_double
    lib.c4ref_submat_lookup.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.c4ref_submat_lookup.restype = C.c_int
    return lib


def session(fn, options=()):
    """Run fn(lib) inside an initialised reference context.

    options: extra exonerate flags, e.g. ("--dpmemory", "32", "--gapopen", "-12").
    The reference frees its global state when the session returns, so all work
    with RefModel / RefPair must happen inside fn.  One session per process is
    the safe usage (the reference's ArgumentSets are function-static).
    """
    lib = _load()
    argv = [b"c4ref"] + [o.encode() for o in options]
    arr = (C.c_char_p * (len(argv) + 1))(*argv, None)
    box = {}

    def _cb(_ctx):
        box["result"] = fn(lib)
        return 0

    rc = lib.c4ref_session(len(argv), arr, _CB(_cb), None)
    if rc != 0:
        raise RuntimeError("reference session failed rc=%d" % rc)
    return box.get("result")
