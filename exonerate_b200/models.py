"""ctypes binding of libc4host.so: the host-side C4 model layer
(exonerate_b200/csrc/host/): model name -> closed model -> engine tables.

    model, description = host_model("affine:local")            # DNA x DNA
    model, _ = host_model("protein2genome", query_is_protein=True)
"""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libc4host.so")
_lib = None


class Params(C.Structure):
    """C4_Params (csrc/host/c4host.h): the penalties the reference keeps in its
    ArgumentSets (--gapopen, --gapextend, --codongapopen, ...)."""
    _fields_ = [("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("codon_gap_open", C.c_int32),
                ("codon_gap_extend", C.c_int32), ("frameshift", C.c_int32), ("intron_open", C.c_int32),
                ("min_intron", C.c_int32), ("max_intron", C.c_int32), ("match_max_dna", C.c_int32),
                ("match_max_protein", C.c_int32)]


def load_host_library():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError("libc4host.so is not built (run __graft_entry__.build())")
        lib = C.CDLL(HOST_LIB_PATH)
        lib.c4b_host_model.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(Params), C.POINTER(abi.Model),
                                       C.POINTER(C.c_void_p)]
        lib.c4b_host_model.restype = C.c_int
        lib.C4_Params_default.argtypes = [C.POINTER(Params)]
        lib.C4_host_error.restype = C.c_char_p
        lib.c4b_host_splice_arrays.argtypes = [C.c_void_p, C.c_int32, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def default_params():
    p = Params()
    load_host_library().C4_Params_default(C.byref(p))
    return p


def host_model(name, query_is_protein=False, target_is_protein=False, params=None):
    """Build the named shipped model, close it and flatten it.
    Returns (abi.Model, text description of the closed model)."""
    lib = load_host_library()
    m = abi.Model()
    desc = C.c_void_p()
    rc = lib.c4b_host_model(name.encode(), int(query_is_protein), int(target_is_protein),
                            C.byref(params) if params is not None else None, C.byref(m), C.byref(desc))
    text = ""
    if desc.value:
        text = C.string_at(desc.value).decode()
        C.CDLL(None).free(desc)
    if rc == -2:
        raise ValueError("unknown model %r" % name)
    if rc != 0:
        raise RuntimeError("model %r: %s" % (name, lib.C4_host_error().decode()))
    return m, text


def splice_arrays(target, force_gtag=False):
    """The four per-position splice-site score arrays of a genomic sequence
    (SplicePredictor_predict_array_int): list of int32 numpy arrays in
    C4B_SPLICE_* order, ready for PairSet(splice=...)."""
    import numpy as np
    raw = target.encode() if isinstance(target, str) else bytes(target)
    seq = np.frombuffer(raw, dtype=np.uint8)
    out = np.zeros((4, len(seq)), dtype=np.int32)
    load_host_library().c4b_host_splice_arrays(seq.ctypes.data, len(seq), int(force_gtag), out.ctypes.data)
    return [out[k] for k in range(4)]
