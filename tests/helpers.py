"""Shared test infrastructure: golden-fixture loaders, the oracle binding and
seeded synthetic sequence generators.

Everything that touches oracle/ lives here or in tests/ -- never in the product
package.
"""
import ctypes as C
import json
import os
import random
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from exonerate_b200 import abi  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_SO = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")

# ---------------------------------------------------------------------------
# closed-model dumps (written by oracle/ref_driver.c:c4ref_model_dump)
# ---------------------------------------------------------------------------
_KV = re.compile(r'(\w+)=("([^"]*)"|\S*)')


def _parse_line(line):
    kind, _, rest = line.partition(" ")
    d = {}
    for m in _KV.finditer(rest):
        d[m.group(1)] = m.group(3) if m.group(3) is not None else m.group(2)
    return kind, d


def _ints(s):
    return [int(x) for x in s.split(",") if x != ""]


_SPLICE_SITE = {"ss5_forward": abi.SPLICE_5_FORWARD, "ss3_forward": abi.SPLICE_3_FORWARD,
                "ss5_reverse": abi.SPLICE_5_REVERSE, "ss3_reverse": abi.SPLICE_3_REVERSE}


def classify_calc(name, macro, params):
    """Map a reference C4_Calc (name + calc_macro text) to (kind, param[4]).

    This mirrors what the host-side model compiler does by recognising the
    reference's calc identities."""
    p = [0, 0, 0, 0]
    if "split_score_func" in macro:
        kind = abi.CALC_PHASE1_POST if "curr_intron_start >= 1" in macro else abi.CALC_PHASE2_POST
        return kind, p
    if "SplicePrediction_get" in macro:
        site = re.search(r"sps->(ss[35]_(?:forward|reverse))", macro).group(1)
        p[1] = _SPLICE_SITE[site]
        if "intron_open_penalty" in macro:
            p[0] = params["intron_open"]
            return abi.CALC_SPLICE_PRE, p
        return abi.CALC_SPLICE_POST, p
    if "Submat_lookup" in macro:
        if "dna_submat" in macro:
            return abi.CALC_MATCH_DNA, p
        tq = "Sequence_get_symbol(ud->query, %QP+1)" in macro
        tt = "Sequence_get_symbol(ud->target, %TP+1)" in macro
        if tq and tt:
            return abi.CALC_MATCH_3_3, p
        if tt:
            return abi.CALC_MATCH_1_3, p
        if tq:
            return abi.CALC_MATCH_3_1, p
        return abi.CALC_MATCH_PROTEIN, p
    for key in ("codon_gap_open", "codon_gap_extend", "gap_open", "gap_extend"):
        if "aas->" + key + ")" in macro:
            p[0] = params[key]
            return abi.CALC_CONST, p
    if "frameshift_penalty" in macro:
        p[0] = params["frameshift"]
        return abi.CALC_CONST, p
    if name == "match" and macro == "":
        # codon:codon match has no macro in the reference (match.c:543-570 "#if 0")
        return abi.CALC_MATCH_3_3, p
    raise ValueError("unrecognised calc %r macro=%r" % (name, macro))


def parse_model_dump(text, params):
    """dump text -> (abi.Model, info dict with names)."""
    m = abi.Model()
    info = {"states": {}, "transitions": {}, "calcs": {}, "name": None}
    shadows = []
    calc_lines = []
    for line in text.strip().splitlines():
        kind, d = _parse_line(line)
        if kind == "model":
            info["name"] = d["name"]
            m.n_states = int(d["states"])
            m.n_transitions = int(d["transitions"])
            m.n_calcs = int(d["calcs"])
            m.n_shadow_slots = int(d["shadow_designations"])
            m.start_state = int(d["start_state"])
            m.end_state = int(d["end_state"])
            m.start_scope = int(d["start_scope"])
            m.end_scope = int(d["end_scope"])
            m.max_query_advance = int(d["max_query_advance"])
            m.max_target_advance = int(d["max_target_advance"])
            assert d["start_cell_func"] == "0" and d["end_cell_func"] == "0"
        elif kind == "state":
            info["states"][int(d["id"])] = d["name"]
        elif kind == "calc":
            calc_lines.append(d)
        elif kind == "transition":
            t = m.transitions[int(d["id"])]
            t.input, t.output = int(d["input"]), int(d["output"])
            t.advance_query, t.advance_target = int(d["advance_query"]), int(d["advance_target"])
            t.calc, t.label = int(d["calc"]), int(d["label"])
            info["transitions"][int(d["id"])] = dict(d, dst_shadows=_ints(d["dst_shadows"]))
        elif kind == "shadow":
            shadows.append(d)
    # shadows: slot stamping per source state; the slot a POST calc reads
    shadow_slot = {}
    for sh in shadows:
        slot = int(sh["designation"])
        shadow_slot[int(sh["id"])] = slot
        start_kind = 1 if "%TP" in sh["start_macro"] else 2
        for st in _ints(sh["src_states"]):
            m.shadow_start[st][slot] = start_kind
    for d in calc_lines:
        cid = int(d["id"])
        kind, p = classify_calc(d["name"], d["macro"], params)
        c = m.calcs[cid]
        c.kind, c.protect = kind, int(d["protect"])
        if kind in (abi.CALC_SPLICE_POST, abi.CALC_PHASE1_POST, abi.CALC_PHASE2_POST):
            slots = set()
            for tid, tr in info["transitions"].items():
                if int(tr["calc"]) == cid:
                    for sid in tr["dst_shadows"]:
                        slots.add(shadow_slot[sid])
            assert len(slots) == 1, "calc reads more than one slot"
            p[2] = slots.pop()
        for i in range(4):
            c.param[i] = p[i]
        info["calcs"][cid] = d["name"]
    return m, info


def load_params():
    with open(os.path.join(GOLDEN, "scoring.json")) as f:
        return json.load(f)


def load_scoring(params=None):
    params = params or load_params()
    s = abi.Scoring()
    for i, v in enumerate(params["dna_matrix"]):
        s.dna_matrix[i] = v
    for i, v in enumerate(params["protein_matrix"]):
        s.protein_matrix[i] = v
    for name in ("dna_index", "protein_index", "nt2d", "codon_aa"):
        raw = bytes.fromhex(params[name])
        arr = getattr(s, name)
        for i, v in enumerate(raw):
            arr[i] = v
    s.min_intron = params["min_intron"]
    s.max_intron = params["max_intron"]
    return s


def load_model(name, params=None):
    """name e.g. 'affine_local_dna' -> (abi.Model, info)."""
    params = params or load_params()
    with open(os.path.join(GOLDEN, "models", name + ".txt")) as f:
        return parse_model_dump(f.read(), params)


def load_cases(name):
    with open(os.path.join(GOLDEN, "cases_" + name + ".json")) as f:
        return json.load(f)


# ---------------------------------------------------------------------------
# pair marshalling
# ---------------------------------------------------------------------------
class PairBuf:
    """Keeps the numpy buffers a c4b_pair points at alive."""

    def __init__(self, q, t, splice=None, blocked=None, region=None):
        self.q = np.frombuffer(q.encode() if isinstance(q, str) else bytes(q), dtype=np.uint8).copy()
        self.t = np.frombuffer(t.encode() if isinstance(t, str) else bytes(t), dtype=np.uint8).copy()
        self.pair = abi.Pair()
        p = self.pair
        p.query = self.q.ctypes.data
        p.target = self.t.ctypes.data
        p.query_len, p.target_len = len(self.q), len(self.t)
        if region is None:
            region = (0, 0, len(self.q), len(self.t))
        p.query_start, p.target_start, p.query_length, p.target_length = region
        self.splice = None
        if splice is not None:
            self.splice = [np.ascontiguousarray(a, dtype=np.int32) for a in splice]
            for i, a in enumerate(self.splice):
                p.splice[i] = a.ctypes.data
        self.blocked = None
        if blocked:
            pts = sorted(set((tj, qi) for qi, tj in blocked))
            self.blocked = (np.array([qi for tj, qi in pts], dtype=np.int32),
                            np.array([tj for tj, qi in pts], dtype=np.int32))
            p.blocked_query_pos = self.blocked[0].ctypes.data
            p.blocked_target_pos = self.blocked[1].ctypes.data
            p.n_blocked = len(pts)


def result_to_dict(res, ops):
    n = res.n_ops
    o = int(res.ops_offset)
    return {"score": res.score,
            "region": [res.query_start, res.target_start,
                       res.query_end - res.query_start, res.target_end - res.target_start],
            "ops": [(int(ops[2 * (o + i)]), int(ops[2 * (o + i) + 1])) for i in range(n)],
            "status": res.status}


# ---------------------------------------------------------------------------
# the oracle (oracle/c4_oracle.c)
# ---------------------------------------------------------------------------
_oracle = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), ORACLE_SO])


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.c4o_viterbi.argtypes = [C.POINTER(abi.Model), C.POINTER(abi.Scoring), C.POINTER(abi.Pair),
                                    C.c_int, C.POINTER(abi.Result), C.c_void_p, C.c_int64]
        lib.c4o_viterbi.restype = C.c_int
        lib.c4o_find_path.argtypes = [C.POINTER(abi.Model), C.POINTER(abi.Scoring), C.POINTER(abi.Pair),
                                      C.c_int32, C.c_int64, C.POINTER(abi.Result), C.c_void_p, C.c_int64]
        lib.c4o_find_path.restype = C.c_int
        lib.c4o_format_vulgar.argtypes = [C.POINTER(abi.Model), C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        lib.c4o_format_cigar.argtypes = [C.POINTER(abi.Model), C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        lib.c4o_rescore_path.argtypes = [C.POINTER(abi.Model), C.POINTER(abi.Scoring), C.POINTER(abi.Pair),
                                         C.POINTER(abi.Result), C.c_void_p]
        lib.c4o_rescore_path.restype = C.c_int32
        lib.c4o_transition_is_valid.argtypes = [C.POINTER(abi.Model)] + [C.c_int] * 5
        lib.c4o_viterbi_cells.argtypes = [C.POINTER(abi.Model), C.POINTER(abi.Scoring), C.POINTER(abi.Pair),
                                          C.c_int, C.c_void_p, C.c_void_p, C.POINTER(abi.Result), C.c_void_p,
                                          C.c_int64]
        lib.c4o_viterbi_cells.restype = C.c_int
        lib.c4o_hsp_extend_one.argtypes = [C.POINTER(abi.Scoring), C.POINTER(abi.HspParam), C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, abi.HspSeed,
                                           C.POINTER(abi.Hsp)]
        lib.c4o_hsp_extend_one.restype = None
        lib.c4o_hspset_replay.argtypes = [C.POINTER(abi.HspParam), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        lib.c4o_hspset_replay.restype = C.c_int
        _oracle = lib
    return _oracle


def oracle_viterbi(model, scoring, pb, mode, max_ops=1 << 16):
    lib = oracle()
    res = abi.Result()
    ops = np.zeros(2 * max_ops, dtype=np.int32)
    rc = lib.c4o_viterbi(C.byref(model), C.byref(scoring), C.byref(pb.pair), mode, C.byref(res),
                         ops.ctypes.data, max_ops)
    assert rc == 0, "oracle rc=%d" % rc
    return result_to_dict(res, ops)


def oracle_find_path(model, scoring, pb, threshold=abi.IMPOSSIBLY_LOW_SCORE,
                     region_threshold_cells=0, max_ops=1 << 16):
    lib = oracle()
    res = abi.Result()
    ops = np.zeros(2 * max_ops, dtype=np.int32)
    rc = lib.c4o_find_path(C.byref(model), C.byref(scoring), C.byref(pb.pair), threshold,
                           region_threshold_cells, C.byref(res), ops.ctypes.data, max_ops)
    assert rc == 0, "oracle rc=%d" % rc
    return result_to_dict(res, ops)


def format_ops(model, ops, which="vulgar"):
    lib = oracle()
    arr = np.array([x for op in ops for x in op], dtype=np.int32)
    buf = C.create_string_buffer(64 + 24 * max(1, len(ops)))
    fn = lib.c4o_format_vulgar if which == "vulgar" else lib.c4o_format_cigar
    n = fn(C.byref(model), arr.ctypes.data, len(ops), buf, len(buf))
    assert n >= 0
    return buf.value.decode()


def report_line(kind, model, qid, tid, qstrand, tstrand, r, q_len=None, t_len=None):
    """'vulgar: ...' / 'cigar: ...' line exactly as Alignment_display_* prints it
    (alignment.c:1622-1639,2671-2706) for forward-strand sequences."""
    qs, ts, ql, tl = r["region"]
    head = "%s: %s %d %d %s %s %d %d %s %d" % (kind, qid, qs, qs + ql, qstrand, tid, ts, ts + tl,
                                               tstrand, r["score"])
    body = format_ops(model, r["ops"], kind)
    # the separating blank is unconditional; cigar's second blank before the
    # first block comes out of the block printer itself (zero-length head op)
    return head + " " + body


# ---------------------------------------------------------------------------
# seeded synthetic inputs
# ---------------------------------------------------------------------------
def rand_dna(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, rate, alphabet="ACGT"):
    out = []
    for ch in s:
        if rng.random() < rate:
            kind = rng.randrange(3)
            if kind == 0:
                continue  # deletion
            if kind == 1:
                out.append(rng.choice(alphabet))  # insertion
                out.append(ch)
            else:
                out.append(rng.choice(alphabet))  # substitution
        else:
            out.append(ch)
    return "".join(out)


def dna_pair(seed, qlen, tlen, rate=0.15):
    """query qlen bp; target = mutated query planted in random flank, tlen bp."""
    rng = random.Random(seed)
    q = rand_dna(rng, qlen)
    core = mutate(rng, q, rate)
    if len(core) >= tlen:
        return q, core[:tlen]
    off = rng.randrange(0, tlen - len(core) + 1)
    t = rand_dna(rng, off) + core + rand_dna(rng, tlen - len(core) - off)
    return q, t


PROTEIN_ALPHABET = "ARNDCQEGHILKMFPSTWYV"


def protein_pair(seed, qlen, tlen, rate=0.2):
    rng = random.Random(seed)
    q = rand_dna(rng, qlen, PROTEIN_ALPHABET)
    core = mutate(rng, q, rate, PROTEIN_ALPHABET)
    if len(core) >= tlen:
        return q, core[:tlen]
    off = rng.randrange(0, tlen - len(core) + 1)
    t = rand_dna(rng, off, PROTEIN_ALPHABET) + core + rand_dna(rng, tlen - len(core) - off, PROTEIN_ALPHABET)
    return q, t


def gene_pair(seed, qlen, tlen, n_exons=5, rate=0.02, reverse=False):
    """cDNA query of n_exons exons; target = the exons (lightly mutated) separated
    by GT..AG introns (CT..AC for a reverse-strand gene) planted in random flank."""
    rng = random.Random(seed)
    n_exons = max(1, min(n_exons, qlen // 8 or 1))
    cuts = sorted(rng.sample(range(1, qlen), n_exons - 1)) if n_exons > 1 else []
    exons = [e for e in (("x" * qlen)[a:b] for a, b in zip([0] + cuts, cuts + [qlen]))]
    q = rand_dna(rng, qlen)
    pos, parts = 0, []
    for e in exons:
        parts.append(q[pos:pos + len(e)])
        pos += len(e)
    spare = max(0, tlen * 3 // 4 - qlen)
    intron = max(35, spare // max(1, n_exons - 1))
    donor, acceptor = ("CT", "AC") if reverse else ("GT", "AG")
    body = ""
    for k, e in enumerate(parts):
        body += mutate(rng, e, rate)
        if k + 1 < len(parts):
            body += donor + rand_dna(rng, max(1, intron - 4)) + acceptor
    if len(body) >= tlen:
        return q, body[:tlen]
    left = rng.randrange(0, tlen - len(body) + 1)
    return q, rand_dna(rng, left) + body + rand_dna(rng, tlen - len(body) - left)


# ---------------------------------------------------------------------------
# HSP seeding / extension
# ---------------------------------------------------------------------------
HSP_MATCH_KIND = {"dna2dna": abi.CALC_MATCH_DNA, "protein2protein": abi.CALC_MATCH_PROTEIN,
                  "protein2dna": abi.CALC_MATCH_1_3, "dna2protein": abi.CALC_MATCH_3_1,
                  "codon2codon": abi.CALC_MATCH_3_3}


def hsp_param(case):
    p = case["params"]
    return abi.HspParam(HSP_MATCH_KIND[case["match"]], p["seedlen"], p["dropoff"], p["threshold"])


def softmask_bytes(seq, enabled):
    """Alphabet_is_masked for a soft-masked alphabet: lower case (alphabet.h:87-88)."""
    if not enabled:
        return None
    return np.array([1 if c.islower() else 0 for c in seq], dtype=np.uint8)


def hsp_tuple(h):
    return [h.query_start, h.target_start, h.length, h.score, h.cobs]


def oracle_hsp_extend(scoring, param, q, t, seeds, qmask=None, tmask=None):
    """per-seed results of the oracle (c4o_hsp_extend_one), as abi.Hsp array"""
    lib = oracle()
    qb = np.frombuffer(q.encode(), dtype=np.uint8).copy()
    tb = np.frombuffer(t.encode(), dtype=np.uint8).copy()
    out = (abi.Hsp * max(1, len(seeds)))()
    for k, (qs, ts) in enumerate(seeds):
        lib.c4o_hsp_extend_one(C.byref(scoring), C.byref(param), qb.ctypes.data, len(qb),
                               qmask.ctypes.data if qmask is not None else None, tb.ctypes.data, len(tb),
                               tmask.ctypes.data if tmask is not None else None, abi.HspSeed(qs, ts),
                               C.byref(out[k]))
    return out


def hspset_replay(param, qlen, seeds, ext):
    """HSPset_seed_hsp's diagonal horizon over per-seed results -> stored HSPs (seed order)"""
    lib = oracle()
    n = len(seeds)
    sd = (abi.HspSeed * max(1, n))(*[abi.HspSeed(a, b) for a, b in seeds])
    out = (abi.Hsp * max(1, n))()
    cnt = lib.c4o_hspset_replay(C.byref(param), qlen, n, sd, ext, out)
    return [hsp_tuple(out[k]) for k in range(cnt)]


def oracle_viterbi_cells(model, scoring, pb, mode, start_cells=None, end_cells=None, max_ops=1 << 14):
    """c4o_viterbi_cells: the fill with cell_start_func / cell_end_func as tables"""
    lib = oracle()
    res = abi.Result()
    ops = np.zeros(2 * max_ops, dtype=np.int32)
    rc = lib.c4o_viterbi_cells(C.byref(model), C.byref(scoring), C.byref(pb.pair), mode,
                               start_cells.ctypes.data if start_cells is not None else None,
                               end_cells.ctypes.data if end_cells is not None else None,
                               C.byref(res), ops.ctypes.data, max_ops)
    assert rc == 0, rc
    return result_to_dict(res, ops)
