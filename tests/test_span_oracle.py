"""CPU tier: the oracle's restatement of BSDP span integration (oracle/c4_oracle.c:
c4o_span_integrate, after Heuristic_Span_integrate, src/bsdp/heuristic.c:589-678) against
golden vectors produced by the unmodified reference function
(tests/golden/make_span_golden.py) -- SURVEY.md 8f row 1; the device kernel is next round's."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers

CASES = json.load(open(os.path.join(helpers.GOLDEN, "span_cases.json")))


def oracle_span_integrate(case):
    lib = helpers.oracle()
    I = C.POINTER(C.c_int32)
    lib.c4o_span_integrate.argtypes = [I, I, I, I, I]
    lib.c4o_span_integrate.restype = None
    sc = np.asarray(case["src_scores"], dtype=np.int32)
    src = np.asarray(case["src_region"], dtype=np.int32)
    dst = np.asarray(case["dst_region"], dtype=np.int32)
    sp = np.asarray(case["span"], dtype=np.int32)
    out = np.full(2 * (dst[2] + 1) * (dst[3] + 1), -7, dtype=np.int32)
    as_p = lambda a: a.ctypes.data_as(I)
    lib.c4o_span_integrate(as_p(sc), as_p(src), as_p(dst), as_p(sp), as_p(out))
    return out.tolist()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_span_integrate_vs_reference(case):
    assert oracle_span_integrate(case) == case["positions"]


def test_span_golden_covers_the_interesting_cases():
    by = {c["name"]: c for c in CASES}
    assert all(p == -1 for p in by["window_misses_random"]["positions"])       # empty windows
    assert any(p >= 0 for p in by["intron_adjacent_sparse"]["positions"])
    # ties: the first maximum in (query, target) scan order wins (strict '<', heuristic.c:638)
    c = by["both_axes_ties"]
    sq, st, ql, tl = c["src_region"]
    dq, dt, dql, dtl = c["dst_region"]
    mnq, mxq, mnt, mxt = c["span"]
    sc = np.asarray(c["src_scores"]).reshape(ql + 1, tl + 1)
    checked = 0
    for i in range(dql + 1):
        for j in range(dtl + 1):
            q, t = c["positions"][2 * (i * (dtl + 1) + j):2 * (i * (dtl + 1) + j) + 2]
            x0, x1 = max(sq, dq + i - mxq), min(sq + ql, dq + i - mnq)
            y0, y1 = max(st, dt + j - mxt), min(st + tl, dt + j - mnt)
            if x0 > x1 or y0 > y1:
                assert (q, t) == (-1, -1)
                continue
            win = sc[x0 - sq:x1 - sq + 1, y0 - st:y1 - st + 1]
            first = np.argwhere(win == win.max())[0]          # row-major first
            assert (q, t) == (x0 + first[0], y0 + first[1])
            checked += 1
    assert checked > 50
