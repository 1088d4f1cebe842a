"""CPU tier: the C-ABI library builds, loads and exports every symbol that
include/c4b200.h declares; no compute call is made (no GPU here)."""
import ctypes as C
import os
import re

import pytest

import helpers
from exonerate_b200 import abi, engine


def declared_symbols():
    text = open(os.path.join(helpers.ROOT, "include", "c4b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(c4b_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(engine.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return C.CDLL(engine.LIB_PATH)


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libc4b200.so does not export %s" % s
    assert sorted(engine.EXPORTS) == syms


def test_abi_version(lib):
    lib.c4b_abi_version.restype = C.c_int
    assert lib.c4b_abi_version() == abi.ABI_VERSION


def test_struct_sizes_match_header():
    # sizes the C compiler gives the same declarations (LP64)
    assert C.sizeof(abi.Calc) == 24
    assert C.sizeof(abi.Transition) == 24
    assert C.sizeof(abi.Model) == 40 + 32 * 4 + 64 * 24 + 32 * 24
    assert C.sizeof(abi.Scoring) == 2 * 576 * 4 + 3 * 256 + 4096 + 8
    assert C.sizeof(abi.Pair) == 16 + 24 + 32 + 16 + 8
    assert C.sizeof(abi.Result) == 40
    assert C.sizeof(abi.HspParam) == 16 and C.sizeof(abi.HspSeed) == 8 and C.sizeof(abi.Hsp) == 32


def test_struct_sizes_match_the_c_compiler(tmp_path):
    """compile include/c4b200.h with gcc and compare every sizeof with the ctypes mirror"""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "c4b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(c4b_calc),sizeof(c4b_transition),sizeof(c4b_model),sizeof(c4b_scoring),sizeof(c4b_pair),'
                   'sizeof(c4b_result),sizeof(c4b_hsp_param),sizeof(c4b_hsp_seed),sizeof(c4b_hsp));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(helpers.ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(t) for t in (abi.Calc, abi.Transition, abi.Model, abi.Scoring, abi.Pair, abi.Result,
                                  abi.HspParam, abi.HspSeed, abi.Hsp)]
    assert got == want


def test_no_cpu_fallback(lib):
    """Without a CUDA device the engine must refuse to start, loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib.c4b_engine_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.c4b_last_error.restype = C.c_char_p
    h = C.c_void_p()
    assert lib.c4b_engine_create(0, C.byref(h)) != 0
    assert b"no CPU fallback" in lib.c4b_last_error()
    with pytest.raises(engine.C4BError):
        engine.Engine(0)


@pytest.mark.parametrize("name,q_protein", [("affine:local", False), ("est2genome", False),
                                            ("protein2genome", True), ("coding2coding", False),
                                            ("ungapped", False)])
def test_every_shipped_model_specialises_for_sm100a(name, q_protein):
    """c4b_model_specialise: the run-time generated kernel of each shipped model compiles
    (NVRTC, sm_100a) in all three fill modes -- the device analogue of the reference's
    bootstrapper building its model archive.  Compile only: no GPU needed."""
    from exonerate_b200 import load_library
    from exonerate_b200.models import host_model
    lib = load_library()
    model, _ = host_model(name, query_is_protein=q_protein)
    for mode in (0, 1, 2):
        size = C.c_int64(0)
        rc = lib.c4b_model_specialise(C.byref(model), mode, 256, C.byref(size))
        assert rc == 0, lib.c4b_last_error().decode()[:2000]
        assert size.value > 10000
    assert lib.c4b_model_specialise(C.byref(model), 7, 256, None) == -1
    assert lib.c4b_model_specialise(C.byref(model), 0, 100, None) == -1


def test_every_scope_combination_specialises():
    """BSDP derives sub-models with every start / end scope (C4_DerivedModel_create,
    heuristic.c:242-325,445-528): the specialised kernel has a compile-time branch per scope
    and per REGION start-slot layout, so compile them all (est2genome: shadows + splice calcs)."""
    from exonerate_b200 import load_library
    from exonerate_b200.models import host_model
    lib = load_library()
    base, _ = host_model("est2genome")
    combos = [(s, abi.SCOPE_ANYWHERE) for s in range(5)] + [(abi.SCOPE_ANYWHERE, e) for e in range(1, 5)] + \
             [(abi.SCOPE_CORNER, abi.SCOPE_CORNER)]
    for start, end in combos:
        m = type(base).from_buffer_copy(base)
        m.start_scope, m.end_scope = start, end
        for mode in (0, 2):  # FIND_PATH differs from FIND_SCORE only by its traceback stores
            rc = lib.c4b_model_specialise(C.byref(m), mode, 128, None)
            assert rc == 0, (start, end, mode, lib.c4b_last_error().decode()[:1500])


def test_path_pass_variants_compile_for_the_spliced_models():
    """FIND_PATH of the systolic specialisation has three forms -- whole lattice, one column window started
    from a checkpoint (JIT_SYS_WIN), SubOpt blocked cells as per-strip entries (JIT_SYS_BLK) -- and all of
    them must compile for the models with shadows, codon calcs and 2- / 3-column advances (protein2genome,
    coding2coding); est2genome's are compiled by the scope test above.  nvcc never sees this source."""
    from exonerate_b200 import load_library
    from exonerate_b200.models import host_model
    lib = load_library()
    for name, protein in (("protein2genome", True), ("coding2coding", False)):
        model, _ = host_model(name, query_is_protein=protein)
        size = C.c_int64(0)
        assert lib.c4b_model_specialise(C.byref(model), 1, 128, C.byref(size)) == 0, lib.c4b_last_error().decode()[:800]
        assert size.value > 10000


def test_specialise_fills_the_disk_cache(tmp_path, monkeypatch):
    """With C4B_JIT_CACHE_DIR set, c4b_model_specialise leaves every variant's cubin in the cache
    (thread-per-row kernel: ring in shared memory / L2, and both start-slot layouts for FIND_REGION;
    plus the systolic kernel of the mode, its SubOpt form and its column-window variant: the
    checkpointing score pass for FIND_SCORE, the one-window pass for FIND_PATH) -- the device analogue of running the reference's
    bootstrapper once -- and a second call rewrites the same files (names are a hash of the
    generated source)."""
    from exonerate_b200 import load_library
    from exonerate_b200.models import host_model
    lib = load_library()
    monkeypatch.setenv("C4B_JIT_CACHE_DIR", str(tmp_path))
    model, _ = host_model("coding2coding")
    assert lib.c4b_model_specialise(C.byref(model), 0, 128, None) == 0
    first = sorted(os.listdir(tmp_path))
    assert len(first) == 5 and all(f.startswith("c4bjit_") and f.endswith(".cubin") for f in first)
    assert lib.c4b_model_specialise(C.byref(model), 2, 128, None) == 0
    assert len(os.listdir(tmp_path)) == 5 + 6
    assert lib.c4b_model_specialise(C.byref(model), 0, 128, None) == 0
    assert len(os.listdir(tmp_path)) == 11
    assert lib.c4b_model_specialise(C.byref(model), 1, 128, None) == 0
    assert len(os.listdir(tmp_path)) == 11 + 5
    monkeypatch.setenv("C4B_JIT_SYSTOLIC", "0")   # the thread-per-row variants alone
    other = tmp_path / "no_systolic"
    other.mkdir()
    monkeypatch.setenv("C4B_JIT_CACHE_DIR", str(other))
    assert lib.c4b_model_specialise(C.byref(model), 0, 128, None) == 0
    assert len(os.listdir(other)) == 2
    monkeypatch.setenv("C4B_JIT_CACHE_DIR", str(tmp_path))
    assert all(os.path.getsize(tmp_path / f) > 10000 for f in os.listdir(tmp_path) if f.endswith(".cubin"))
