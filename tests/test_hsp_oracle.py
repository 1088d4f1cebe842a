"""CPU tier: the HSP oracle (oracle/c4_oracle.c: c4o_hsp_extend_one + c4o_hspset_replay)
against golden vectors produced by the unmodified reference (HSPset_seed_hsp over a seed
list + HSPset_finalise; tests/golden/make_hsp_golden.py) -- SURVEY.md 8a row a14."""
import json
import os

import pytest

import helpers

CASES = json.load(open(os.path.join(helpers.GOLDEN, "hsp_cases.json")))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_hsp_oracle_vs_reference(case, scoring):
    assert case["params"]["filter_threshold"] == 0 and case["params"]["seed_repeat"] == 1
    param = helpers.hsp_param(case)
    seeds = [tuple(s) for s in case["seeds"]]
    qm = helpers.softmask_bytes(case["q"], case["softmask_query"])
    tm = helpers.softmask_bytes(case["t"], case["softmask_target"])
    ext = helpers.oracle_hsp_extend(scoring, param, case["q"], case["t"], seeds, qm, tm)
    assert all(ext[k].status == 0 for k in range(len(seeds)))
    assert helpers.hspset_replay(param, len(case["q"]), seeds, ext) == case["hsps"]


def test_reference_known_answers():
    """src/comparison/hspset.test.c inputs: the values the reference prints."""
    byname = {c["name"]: c for c in CASES}
    assert byname["ref_test_d2d"]["hsps"] == [[4, 4, 20, 82, 11], [34, 34, 17, 76, 7]]
    assert byname["ref_test_p2d"]["hsps"] == [[5, 15, 12, 49, 6]]
    assert sum(len(c["hsps"]) for c in CASES) >= 150
