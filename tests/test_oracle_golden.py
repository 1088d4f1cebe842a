"""Pins the CPU oracle (oracle/c4_oracle.c) against the reference.

Golden vectors come from running the unmodified reference in the build
container (tests/golden/make_golden.py); they include the reference's own KAT
inputs and asserted scores (src/model/affine.test.c:106-109,
est2genome.test.c:63, protein2genome.test.c:34, coding2coding.test.c:35).
"""
import numpy as np
import pytest

import helpers
from exonerate_b200 import abi

MODELS = ["affine_local_dna", "affine_global_dna", "affine_bestfit_dna", "affine_overlap_dna",
          "affine_local_protein", "affine_global_protein", "affine_bestfit_protein",
          "affine_overlap_protein", "ungapped_dna", "est2genome", "protein2genome", "coding2coding"]

# score assertions in the reference's own tests
KAT_SCORES = {"affine_global_protein": -151, "affine_bestfit_protein": 18,
              "affine_local_protein": 32, "affine_overlap_protein": 18,
              "est2genome": 157, "protein2genome": 125, "coding2coding": 169}


def splice_for(name, case):
    if name not in ("est2genome", "protein2genome"):
        return None
    z = np.load(helpers.GOLDEN + "/splice_%s.npz" % name)
    return [z["%s_%d" % (case["name"], ty)] for ty in range(4)]


def cases_for(name):
    return [pytest.param(name, c, id="%s-%s" % (name, c["name"])) for c in helpers.load_cases(name)]


ALL_CASES = [p for name in MODELS for p in cases_for(name)]


@pytest.mark.parametrize("name", sorted(KAT_SCORES))
def test_reference_kat_scores_in_golden(name):
    """The golden file itself reproduces the reference's asserted KAT scores."""
    kat = [c for c in helpers.load_cases(name) if c["name"] == "kat"][0]
    assert kat["score"] == KAT_SCORES[name]
    assert kat["path"]["score"] == KAT_SCORES[name]


@pytest.mark.parametrize("name,case", ALL_CASES)
def test_oracle_matches_reference(name, case, params, scoring):
    model, _ = helpers.load_model(name, params)
    pb = helpers.PairBuf(case["q"], case["t"], splice=splice_for(name, case))
    ref = case["path"]
    # FIND_SCORE
    r = helpers.oracle_viterbi(model, scoring, pb, abi.MODE_FIND_SCORE)
    assert r["score"] == case["score"]
    # FIND_PATH, quadratic traceback over the whole lattice
    r = helpers.oracle_viterbi(model, scoring, pb, abi.MODE_FIND_PATH)
    assert r["score"] == ref["score"]
    assert r["region"] == ref["region"]
    assert r["ops"] == [tuple(o) for o in ref["ops"]]
    # report strings
    assert helpers.report_line("vulgar", model, "qy", "tg", *strands(name), r) == ref["vulgar"]
    assert helpers.report_line("cigar", model, "qy", "tg", *strands(name), r) == ref["cigar"]
    # FIND_REGION shadows give the same bounding box
    rr = helpers.oracle_viterbi(model, scoring, pb, abi.MODE_FIND_REGION)
    assert rr["score"] == ref["score"]
    assert rr["region"] == ref["region"]
    # Optimal_find_path flow: region pass then path inside the box
    # (EDGE/QUERY/TARGET scopes re-apply to the sub-box in the reference's reduced-space
    # flow, optimal.c:382-399, so that flow is only self-consistent for these scopes)
    if model.start_scope in (abi.SCOPE_ANYWHERE, abi.SCOPE_CORNER):
        r2 = helpers.oracle_find_path(model, scoring, pb, region_threshold_cells=0)
        assert r2["ops"] == r["ops"] and r2["region"] == r["region"] and r2["score"] == r["score"]


def strands(name):
    q = "." if name.endswith("protein") or name == "protein2genome" else "+"
    t = "." if name.endswith("protein") else "+"
    return q, t


@pytest.mark.parametrize("name,case", ALL_CASES)
def test_oracle_rescore(name, case, params, scoring):
    """Alignment_is_valid's consistency check: the path re-scores to the DP score."""
    import ctypes as C
    model, _ = helpers.load_model(name, params)
    pb = helpers.PairBuf(case["q"], case["t"], splice=splice_for(name, case))
    lib = helpers.oracle()
    res = abi.Result()
    ops = np.zeros(2 << 14, dtype=np.int32)
    assert lib.c4o_viterbi(C.byref(model), C.byref(scoring), C.byref(pb.pair), abi.MODE_FIND_PATH,
                           C.byref(res), ops.ctypes.data, 1 << 14) == 0
    s = lib.c4o_rescore_path(C.byref(model), C.byref(scoring), C.byref(pb.pair), C.byref(res),
                             ops.ctypes.data)
    assert s == res.score == case["path"]["score"]


def blocked_points(model, alignments):
    """SubOpt_add_alignment (src/c4/subopt.c:66-140) for advance<=1 models:
    the source cell of every MATCH step of every earlier alignment."""
    pts = []
    for a in alignments:
        qp, tp = a["region"][0], a["region"][1]
        for tid, length in a["ops"]:
            tr = model.transitions[tid]
            for _ in range(length):
                if tr.label == abi.LABEL_MATCH and (qp, tp) not in pts:
                    pts.append((qp, tp))
                qp += tr.advance_query
                tp += tr.advance_target
    return pts


def test_oracle_subopt_series(params, scoring):
    """--subopt: later alignments avoid MATCH cells of earlier ones (a10)."""
    model, _ = helpers.load_model("affine_local_dna", params)
    n = 0
    for case in helpers.load_cases("affine_local_dna"):
        series = case.get("subopt_series")
        if not series:
            continue
        done = []
        for ref in series:
            pb = helpers.PairBuf(case["q"], case["t"], blocked=blocked_points(model, done))
            r = helpers.oracle_viterbi(model, scoring, pb, abi.MODE_FIND_PATH)
            assert r["score"] == ref["score"]
            assert r["region"] == ref["region"]
            assert r["ops"] == [tuple(o) for o in ref["ops"]]
            done.append(ref)
            n += 1
    assert n >= 6


def test_layout_validity_edges(params):
    """Layout masks (src/c4/layout.c): START/END scope rules at lattice edges."""
    import ctypes as C
    lib = helpers.oracle()
    model, info = helpers.load_model("affine_global_dna", params)
    start_t = [k for k in range(model.n_transitions) if model.transitions[k].input == model.start_state][0]
    end_t = [k for k in range(model.n_transitions) if model.transitions[k].output == model.end_state][0]
    Q, T = 5, 7
    for i in range(Q + 1):
        for j in range(T + 1):
            assert bool(lib.c4o_transition_is_valid(C.byref(model), start_t, i, j, Q, T)) == (i == 0 and j == 0)
            assert bool(lib.c4o_transition_is_valid(C.byref(model), end_t, i, j, Q, T)) == (i == Q and j == T)
    model, _ = helpers.load_model("affine_overlap_dna", params)
    for i in range(Q + 1):
        for j in range(T + 1):
            assert bool(lib.c4o_transition_is_valid(C.byref(model), start_t, i, j, Q, T)) == (i == 0 or j == 0)
            assert bool(lib.c4o_transition_is_valid(C.byref(model), end_t, i, j, Q, T)) == (i == Q or j == T)
