"""CPU tier: the host-side C4 model layer (exonerate_b200/csrc/host) closes the
shipped models into EXACTLY the tables the reference's own C4_Model_close
produces -- states, transition precedence order, calcs, shadow slots -- as
dumped from the reference (tests/golden/models/*.txt)."""
import re

import pytest

import helpers
from exonerate_b200 import abi
from exonerate_b200.models import Params, default_params, host_model

# fixture name -> (model name as on the exonerate command line, query protein, target protein)
CASES = {
    "affine_local_dna": ("affine:local", 0, 0), "affine_global_dna": ("affine:global", 0, 0),
    "affine_bestfit_dna": ("affine:bestfit", 0, 0), "affine_overlap_dna": ("affine:overlap", 0, 0),
    "affine_local_protein": ("a:l", 1, 1), "affine_global_protein": ("a:g", 1, 1),
    "affine_bestfit_protein": ("a:b", 1, 1), "affine_overlap_protein": ("a:o", 1, 1),
    "ungapped_dna": ("ungapped", 0, 0), "est2genome": ("est2genome", 0, 0),
    "protein2genome": ("protein2genome", 1, 0), "coding2coding": ("coding2coding", 0, 0),
}


def records(text, kind):
    out = []
    for line in text.strip().splitlines():
        k, _, rest = line.partition(" ")
        if k == kind:
            out.append(dict((m.group(1), m.group(3) if m.group(3) is not None else m.group(2))
                            for m in re.finditer(r'(\w+)=("([^"]*)"|\S*)', rest)))
    return out


@pytest.mark.parametrize("fixture", sorted(CASES))
def test_closed_model_tables_identical(fixture, params):
    name, qp, tp = CASES[fixture]
    got, _ = host_model(name, qp, tp)
    want, _ = helpers.load_model(fixture, params)
    assert bytes(got) == bytes(want)


@pytest.mark.parametrize("fixture", sorted(CASES))
def test_names_and_bookkeeping_match_reference_dump(fixture):
    """Beyond the engine tables: state / transition / shadow names, ids, shadow
    designations, portals and spans equal the reference's closed model."""
    name, qp, tp = CASES[fixture]
    _, mine = host_model(name, qp, tp)
    ref = open(helpers.GOLDEN + "/models/%s.txt" % fixture).read()
    for kind, keys in (("model", ["name", "states", "transitions", "calcs", "shadows", "portals", "spans",
                                  "max_query_advance", "max_target_advance", "shadow_designations",
                                  "start_state", "start_scope", "end_state", "end_scope"]),
                       ("state", ["id", "name", "src_shadows"]),
                       ("transition", ["id", "name", "input", "output", "advance_query", "advance_target",
                                       "calc", "label", "dst_shadows"]),
                       ("shadow", ["id", "name", "designation", "src_states", "dst_transitions"]),
                       ("calc", ["id", "name", "protect"]),
                       ("portal", ["id", "name", "advance_query", "advance_target", "calc"]),
                       ("span", ["id", "name", "state", "min_query", "max_query", "min_target", "max_target"])):
        a, b = records(mine, kind), records(ref, kind)
        assert len(a) == len(b), kind
        for x, y in zip(a, b):
            assert [x[k] for k in keys] == [y[k] for k in keys], (kind, x, y)


def test_penalties_are_parameters():
    """--gapopen / --gapextend etc. reach the calc tables (affine.c:24-35)."""
    p = default_params()
    assert (p.gap_open, p.gap_extend, p.codon_gap_open, p.codon_gap_extend) == (-12, -4, -18, -8)
    p.gap_open, p.gap_extend = -20, -2
    m, _ = host_model("affine:local", params=p)
    consts = sorted(m.calcs[k].param[0] for k in range(m.n_calcs) if m.calcs[k].kind == abi.CALC_CONST)
    assert consts == [-20, -2]
    m, _ = host_model("coding2coding")          # codon penalties when the match advances by 3
    consts = sorted(m.calcs[k].param[0] for k in range(m.n_calcs) if m.calcs[k].kind == abi.CALC_CONST)
    assert consts == [-28, -18, -8]


def test_other_models_and_errors():
    m, text = host_model("protein2dna", query_is_protein=True)
    assert m.n_states == 6 and m.max_target_advance == 3 and "frameshift p2d" in text
    m, _ = host_model("p2g:b", query_is_protein=True)
    assert m.start_scope == abi.SCOPE_QUERY and m.n_shadow_slots == 1
    with pytest.raises(ValueError):
        host_model("no-such-model")


def test_host_tables_drive_the_oracle(params, scoring):
    """End to end on the CPU: host-built tables + oracle reproduce a reference KAT."""
    model, _ = host_model("affine:local", 1, 1)
    kat = [c for c in helpers.load_cases("affine_local_protein") if c["name"] == "kat"][0]
    r = helpers.oracle_viterbi(model, scoring, helpers.PairBuf(kat["q"], kat["t"]), abi.MODE_FIND_PATH)
    assert r["score"] == 32 and r["ops"] == [tuple(o) for o in kat["path"]["ops"]]


@pytest.mark.parametrize("name", ["est2genome", "protein2genome"])
def test_splice_arrays_match_reference(name):
    """Host splice predictor == SplicePredictor_predict_array_int of the reference
    (float32 PSSM log-odds, windowed sums clipped at the ends, rounding)."""
    import numpy as np
    from exonerate_b200.models import splice_arrays
    z = np.load(helpers.GOLDEN + "/splice_%s.npz" % name)
    n = 0
    for case in helpers.load_cases(name):
        mine = splice_arrays(case["t"])
        for ty in range(4):
            assert (mine[ty] == z["%s_%d" % (case["name"], ty)]).all(), (case["name"], ty)
            n += 1
    assert n >= 4
