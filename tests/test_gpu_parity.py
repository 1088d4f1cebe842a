"""GPU tier: the CUDA path (through the C ABI) against the oracle and the
committed golden vectors.  Bit-exact: scores, alignment regions, op lists and
the vulgar / cigar strings built from them."""
import random

import numpy as np
import pytest

import helpers
from exonerate_b200 import abi

pytestmark = pytest.mark.gpu

AFFINE = ["affine_local_dna", "affine_global_dna", "affine_bestfit_dna", "affine_overlap_dna",
          "affine_local_protein", "affine_global_protein", "affine_bestfit_protein",
          "affine_overlap_protein"]
GENERIC = ["ungapped_dna", "est2genome", "protein2genome", "coding2coding"]


@pytest.fixture(scope="module")
def eng():
    from exonerate_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def splice_for(name, case):
    if name not in ("est2genome", "protein2genome"):
        return None
    z = np.load(helpers.GOLDEN + "/splice_%s.npz" % name)
    return [z["%s_%d" % (case["name"], ty)] for ty in range(4)]


def strands(name):
    q = "." if name.endswith("protein") or name == "protein2genome" else "+"
    t = "." if name.endswith("protein") else "+"
    return q, t


@pytest.mark.parametrize("name", AFFINE + GENERIC)
def test_golden_vectors(eng, name, params, scoring):
    """Every golden case of the reference, one batch per model."""
    from exonerate_b200 import Batch, Optimal, PairSet
    model, _ = helpers.load_model(name, params)
    cases = helpers.load_cases(name)
    pairs = PairSet([c["q"] for c in cases], [c["t"] for c in cases],
                    splice=[splice_for(name, c) for c in cases])
    b = Batch(eng, model, scoring, pairs, want_path=True)
    want_kernel = "affine_systolic" if name in AFFINE else \
        "e2g_packed16" if name == "est2genome" else "generic_wavefront"
    assert b.kernel_name == want_kernel
    b.close()
    opt = Optimal(eng, model, scoring)
    scores = opt.find_score(pairs)
    paths = opt.find_path(pairs)
    for c, s, r in zip(cases, scores, paths):
        ref = c["path"]
        assert s == c["score"], c["name"]
        assert r["score"] == ref["score"], c["name"]
        assert r["region"] == ref["region"], c["name"]
        assert r["ops"] == [tuple(o) for o in ref["ops"]], c["name"]
        assert helpers.report_line("vulgar", model, "qy", "tg", *strands(name), r) == ref["vulgar"]
        assert helpers.report_line("cigar", model, "qy", "tg", *strands(name), r) == ref["cigar"]


@pytest.mark.parametrize("systolic", ["1", "0"])
@pytest.mark.parametrize("name", AFFINE + GENERIC)
def test_golden_vectors_specialised_kernel(eng, name, params, scoring, monkeypatch, systolic):
    """The same golden cases through the run-time SPECIALISED table-driven kernels, compiled
    for this model by NVRTC: the systolic one (generic_jit_systolic.cuh: lattice in registers,
    lanes skewed by a column) and the thread-per-row one (generic_jit_kernel.cuh); every model,
    the affine and est2genome families included, forced off their own kernels."""
    from exonerate_b200 import Batch, Optimal, PairSet
    monkeypatch.setenv("C4B_FORCE_GENERIC", "1")
    monkeypatch.setenv("C4B_GENERIC_JIT", "1")
    monkeypatch.setenv("C4B_JIT_SYSTOLIC", systolic)
    model, _ = helpers.load_model(name, params)
    cases = helpers.load_cases(name)
    pairs = PairSet([c["q"] for c in cases], [c["t"] for c in cases],
                    splice=[splice_for(name, c) for c in cases])
    b = Batch(eng, model, scoring, pairs, want_path=True)
    b.run()
    assert b.kernel_name == ("generic_jit_systolic" if systolic == "1" else "generic_jit")
    b.close()
    opt = Optimal(eng, model, scoring)
    scores = opt.find_score(pairs)
    paths = opt.find_path(pairs)
    for c, s, r in zip(cases, scores, paths):
        ref = c["path"]
        assert s == c["score"], c["name"]
        assert r["score"] == ref["score"], c["name"]
        assert r["region"] == ref["region"], c["name"]
        assert r["ops"] == [tuple(o) for o in ref["ops"]], c["name"]
        assert helpers.report_line("vulgar", model, "qy", "tg", *strands(name), r) == ref["vulgar"]


@pytest.mark.parametrize("wcols", ["32", "128", "1024"])
def test_windowed_path_on_the_table_driven_kernel(eng, params, scoring, monkeypatch, wcols):
    """PATH records that do not fit the device budget (forced here with a 1 kB budget): the systolic
    specialisation leaves column checkpoints of its register lattice and the traceback refills one
    window at a time under the cursor -- the reference recurses through checkpoint rows for the
    same reason (optimal.c:183-345) and results do not depend on it.  Golden cases of every model
    (op for op against the reference) and larger spliced lattices against the whole-record pass:
    several strips, windows narrower than an intron, ANYWHERE scopes (REGION + box) and global ones."""
    from exonerate_b200 import Batch, Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    monkeypatch.setenv("C4B_FORCE_GENERIC", "1")
    monkeypatch.setenv("C4B_GENERIC_JIT", "1")
    monkeypatch.setenv("C4B_GENERIC_WINDOW_COLS", wcols)

    def both(model, pairs):
        opt = Optimal(eng, model, scoring)
        monkeypatch.delenv("C4B_GENERIC_TB_BUDGET_KB", raising=False)
        whole = opt.find_path(pairs)
        monkeypatch.setenv("C4B_GENERIC_TB_BUDGET_KB", "1")
        b = Batch(eng, model, scoring, pairs, want_path=True)
        b.run()
        assert b.kernel_name == "generic_jit_systolic"
        b.close()
        return whole, opt.find_path(pairs)

    if wcols == "128":
        for name in ("affine_local_dna", "affine_global_dna", "est2genome", "protein2genome", "coding2coding"):
            model, _ = helpers.load_model(name, params)
            cases = helpers.load_cases(name)
            pairs = PairSet([c["q"] for c in cases], [c["t"] for c in cases], splice=[splice_for(name, c) for c in cases])
            whole, win = both(model, pairs)
            for c, w, r in zip(cases, whole, win):
                assert r == w, (name, c["name"])
                assert r["score"] == c["path"]["score"] and r["ops"] == [tuple(o) for o in c["path"]["ops"]], (name, c["name"])
    rng = random.Random(91)
    for name in ("protein2genome", "est2genome", "coding2coding", "affine_global_dna"):
        model, _ = helpers.load_model(name, params)
        qs, ts, sp = [], [], []
        for k in range(4):
            if name == "protein2genome":
                q, t = bench_p2g_pair(7000 + k, rng.choice([120, 330]), rng.choice([5000, 9000]))
            elif name == "est2genome":
                q, t = helpers.gene_pair(7100 + k, rng.choice([300, 700]), rng.choice([6000, 12000]), n_exons=3, rate=0.02,
                                         reverse=bool(k & 1))
            else:
                q, t = helpers.dna_pair(7200 + k, rng.choice([200, 450]), rng.choice([1500, 3000]))
            qs.append(q); ts.append(t)
            sp.append(splice_arrays(t) if name in ("protein2genome", "est2genome") else None)
        whole, win = both(model, PairSet(qs, ts, splice=sp))
        assert any(len(r["ops"]) > 3 for r in whole), name
        for k, (w, r) in enumerate(zip(whole, win)):
            assert r == w, (name, k, wcols)


def bench_p2g_pair(seed, qlen, tlen):
    """one pair of bench.py's protein2genome generator (a planted spliced gene)"""
    import bench
    q, t = bench.make_batch_p2g(seed, 1, qlen, tlen)
    return bytes(q[0]).decode(), bytes(t[0]).decode()


def test_specialised_kernel_matches_interpreter_at_size(eng, params, scoring, monkeypatch):
    """protein2genome and coding2coding on lattices with several rows per thread
    (query longer than the 512-thread CTA) and suboptimal-blocked cells: the
    specialised kernel and the interpreter kernel must agree op for op."""
    from exonerate_b200 import Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    rng = random.Random(77)
    for name in ("protein2genome", "coding2coding", "est2genome"):
        model, _ = helpers.load_model(name, params)
        qs, ts, sp = [], [], []
        for k in range(6):
            if name == "protein2genome":
                q = "".join(rng.choice("ACDEFGHIKLMNPQRSTVWY") for _ in range(rng.choice([90, 300, 620])))
                t = helpers.rand_dna(rng, rng.choice([700, 2500, 4000]))
            else:
                q, t = helpers.dna_pair(4400 + k, rng.choice([150, 400, 700]), rng.choice([900, 2000]))
            qs.append(q); ts.append(t)
            sp.append(splice_arrays(t) if name != "coding2coding" else None)
        pairs = PairSet(qs, ts, splice=sp)
        got = {}
        # interpreter, thread-per-row specialisation, systolic specialisation (REGION + box and
        # one PATH pass over the full lattices), systolic with one row per lane (more strips)
        for jit, env in (("0", {}), ("1", {"C4B_JIT_SYSTOLIC": "0"}),
                         ("sys", {"C4B_GENERIC_DIRECT_PATH": "0"}), ("sys-direct", {"C4B_GENERIC_DIRECT_PATH": "1"}),
                         ("sys-r1", {"C4B_JIT_SYS_R": "1"})):
            monkeypatch.setenv("C4B_FORCE_GENERIC", "1")
            monkeypatch.setenv("C4B_GENERIC_JIT", "0" if jit == "0" else "1")
            for k in ("C4B_JIT_SYSTOLIC", "C4B_GENERIC_DIRECT_PATH", "C4B_JIT_SYS_R"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            opt = Optimal(eng, model, scoring)
            got[jit] = (opt.find_score(pairs), opt.find_path(pairs))
        for jit in got:
            assert got["0"][0] == got[jit][0], (name, jit)
            for a, b in zip(got["0"][1], got[jit][1]):
                assert a["score"] == b["score"] and a["region"] == b["region"] and a["ops"] == b["ops"], (name, jit)


def test_est2genome_packed_kernel_vs_specialised_table_driven(eng, params, scoring, monkeypatch):
    """Two independent device implementations of est2genome on lattices too large for the
    oracle to check in seconds (20 genes of 600 bp x 20 kbp): the packed 16-bit kernel with
    checkpoint windows against the run-time specialised table-driven kernel, op for op.  The
    batch is big enough (> 1 MB of sequences + splice arrays) for the direct staging path."""
    from exonerate_b200 import Batch, Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    import bench
    model, _ = helpers.load_model("est2genome", params)
    queries, targets = bench.make_batch_e2g(77, 20, 600, 20000)
    qs = [bytes(q).decode() for q in queries]
    ts = [bytes(t).decode() for t in targets]
    pairs = PairSet(qs, ts, splice=[splice_arrays(t) for t in ts])
    opt = Optimal(eng, model, scoring)
    packed = opt.find_path(pairs)
    monkeypatch.setenv("C4B_NO_E2G", "1")
    monkeypatch.setenv("C4B_GENERIC_JIT", "1")
    b = Batch(eng, model, scoring, pairs, want_path=True)
    b.run()
    assert b.kernel_name == "generic_jit_systolic"
    b.close()
    generic = Optimal(eng, model, scoring).find_path(pairs)
    for k, (a, g) in enumerate(zip(packed, generic)):
        assert a["score"] == g["score"] and a["region"] == g["region"] and a["ops"] == g["ops"], k
        assert a["score"] > 2000  # the planted gene was found


def oracle_path(model, scoring, q, t):
    return helpers.oracle_viterbi(model, scoring, helpers.PairBuf(q, t), abi.MODE_FIND_PATH,
                                  max_ops=len(q) + len(t) + 8)


@pytest.mark.parametrize("name", ["affine_local_dna", "affine_global_dna", "affine_bestfit_dna",
                                  "affine_overlap_dna"])
def test_random_dna_vs_oracle(eng, name, params, scoring):
    """Seeded random lattices incl. ragged sizes around the strip widths
    (R*32 rows) and multi-sweep queries."""
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model(name, params)
    rng = random.Random(11)
    shapes = [(1, 1), (1, 40), (40, 1), (7, 9), (31, 33), (32, 32), (255, 300), (256, 257), (257, 90),
              (511, 600), (512, 512), (513, 700), (1000, 1000), (1023, 1100), (1024, 1500),
              (1025, 1200), (2100, 2300)]
    qs, ts = [], []
    for k, (ql, tl) in enumerate(shapes):
        q, t = helpers.dna_pair(9000 + k, ql, tl, rate=rng.choice([0.05, 0.15, 0.3]))
        qs.append(q)
        ts.append(t)
    # unrelated noise and N-rich inputs
    qs.append(helpers.rand_dna(rng, 400)); ts.append(helpers.rand_dna(rng, 900))
    qs.append(helpers.rand_dna(rng, 300, "ACGTN")); ts.append(helpers.rand_dna(rng, 500, "ACGTNRY"))
    pairs = PairSet(qs, ts)
    opt = Optimal(eng, model, scoring)
    scores = opt.find_score(pairs)
    paths = opt.find_path(pairs)
    for k in range(pairs.n):
        want = oracle_path(model, scoring, qs[k], ts[k])
        assert scores[k] == want["score"], (k, len(qs[k]), len(ts[k]))
        assert paths[k]["score"] == want["score"], k
        assert paths[k]["region"] == want["region"], (k, len(qs[k]), len(ts[k]))
        assert paths[k]["ops"] == want["ops"], (k, len(qs[k]), len(ts[k]))


def test_small_batches_each_strip_width(eng, params, scoring):
    """R = 8 / 16 / 32 kernel instantiations (chosen from the longest query)."""
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    opt = Optimal(eng, model, scoring)
    for maxq in (100, 400, 900):
        qs, ts = [], []
        for k in range(24):
            q, t = helpers.dna_pair(maxq * 100 + k, max(1, maxq - 13 * k % maxq), 150 + 37 * k)
            qs.append(q)
            ts.append(t)
        pairs = PairSet(qs, ts)
        paths = opt.find_path(pairs)
        for k in range(pairs.n):
            want = oracle_path(model, scoring, qs[k], ts[k])
            assert paths[k]["score"] == want["score"] and paths[k]["region"] == want["region"]
            assert paths[k]["ops"] == want["ops"]


def test_banded_two_pass_route(eng, params, scoring):
    """Long targets take score pass -> band refill -> traceback; the result must
    equal the full-lattice traceback of the oracle."""
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    qs, ts = [], []
    for k, (ql, tl) in enumerate([(60, 5000), (200, 8000), (333, 20000), (40, 3000), (1000, 12000)]):
        q, t = helpers.dna_pair(7000 + k, ql, tl)
        qs.append(q)
        ts.append(t)
    # a target with two equally good copies: END tie-break (first in scan order)
    q = helpers.rand_dna(random.Random(5), 80)
    qs.append(q)
    ts.append(helpers.rand_dna(random.Random(6), 1000) + q + helpers.rand_dna(random.Random(7), 2500) + q
              + helpers.rand_dna(random.Random(8), 700))
    pairs = PairSet(qs, ts)
    opt = Optimal(eng, model, scoring)
    paths = opt.find_path(pairs)
    for k in range(pairs.n):
        want = oracle_path(model, scoring, qs[k], ts[k])
        assert paths[k]["score"] == want["score"], k
        assert paths[k]["region"] == want["region"], k
        assert paths[k]["ops"] == want["ops"], k


def test_packed16_score_pass(eng, params, scoring, monkeypatch):
    """affine_fill16u_kernel / affine_fill16_kernel (two lattices per warp in 16-bit
    halves, offset-binary and signed variants) are the score pass of ACGT-only local batches: ragged partners, odd counts, identical and
    unrelated sequences, every strip width, against the oracle; and against the
    int32 kernel (C4B_AFFINE_PACK16=0) on the same batch."""
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    opt = Optimal(eng, model, scoring)
    rng = random.Random(77)
    # (the one-sweep score pass runs the smallest multiple of 4 rows per lane that holds the longest query:
    # 4 .. 32 rows, one instantiation each)
    for maxq in (30, 100, 128, 250, 370, 500, 630, 760, 890, 1023):
        qs, ts = [], []
        shapes = [(maxq, 700), (1, 1), (maxq, 40), (max(1, maxq // 3), 2500), (max(1, maxq - 1), 33)]
        shapes += [(rng.randrange(1, maxq + 1), rng.randrange(1, 1800)) for _ in range(8)]
        for k, (ql, tl) in enumerate(shapes):
            q, t = helpers.dna_pair(maxq * 131 + k, ql, tl, rate=rng.choice([0.0, 0.1, 0.3]))
            qs.append(q)
            ts.append(t)
        same = helpers.rand_dna(rng, maxq)          # the largest reachable score for this Q
        qs.append(same); ts.append(same)
        qs.append("A" * min(maxq, 200)); ts.append("C" * 300)   # best score 0 -> END at (0,0)
        pairs = PairSet(qs, ts)
        assert pairs.n % 2 == 1
        scores = opt.find_score(pairs)
        paths = opt.find_path(pairs)
        monkeypatch.setenv("C4B_AFFINE_PACK16", "0")
        scores32 = opt.find_score(pairs)
        paths32 = opt.find_path(pairs)
        monkeypatch.delenv("C4B_AFFINE_PACK16")
        assert scores == scores32 and paths == paths32
        monkeypatch.setenv("C4B_P16_FOLD", "0")         # two lattices per warp also where a small batch would fold
        from exonerate_b200 import Batch
        b = Batch(eng, model, scoring, pairs, want_path=False)
        need = (maxq + 1 + 31) // 32
        assert "%d rows/lane" % (max(4, (need + 1) // 2 * 2) if need <= 8 else (need + 3) // 4 * 4) in b.description, b.description
        b.close()
        assert opt.find_score(pairs) == scores and opt.find_path(pairs) == paths
        monkeypatch.setenv("C4B_P16_FOLD", "1")         # ... and one lattice per warp, folded (queries of > 255 symbols)
        b = Batch(eng, model, scoring, pairs, want_path=False)
        if maxq > 255:
            assert "folded" in b.description and "%d rows/lane" % max(6, ((maxq + 1 + 63) // 64 + 1) // 2 * 2) in b.description
        b.close()
        assert opt.find_score(pairs) == scores and opt.find_path(pairs) == paths
        monkeypatch.delenv("C4B_P16_FOLD")
        monkeypatch.setenv("C4B_P16_VARIANT", "s")      # signed-halfword variant of the packed kernel
        assert opt.find_score(pairs) == scores and opt.find_path(pairs) == paths
        monkeypatch.delenv("C4B_P16_VARIANT")
        monkeypatch.setenv("C4B_AFFINE_TB16", "0")      # int32 traceback pass under the packed score pass
        assert opt.find_path(pairs) == paths
        monkeypatch.delenv("C4B_AFFINE_TB16")
        for k in range(pairs.n):
            want = oracle_path(model, scoring, qs[k], ts[k])
            assert scores[k] == want["score"], (maxq, k)
            assert paths[k]["region"] == want["region"] and paths[k]["ops"] == want["ops"], (maxq, k)
    # metric shape: the END cell (score, query_end, target_end) of both kernels agrees
    qs, ts = [], []
    for k in range(5):
        q, t = helpers.dna_pair(52000 + k, 1000, 100000 - 1000 * k)
        qs.append(q)
        ts.append(t)
    pairs = PairSet(qs, ts)
    got = opt.find_path(pairs)
    monkeypatch.setenv("C4B_AFFINE_PACK16", "0")
    want = opt.find_path(pairs)
    monkeypatch.delenv("C4B_AFFINE_PACK16")
    assert got == want and all(r["score"] > 2000 for r in got)


def test_staging_pipeline_many_slices(eng, params, scoring, monkeypatch):
    """The staging pipeline (sequence slices copied + encoded on the copy stream,
    score-pass launch groups on the aux streams as their slice lands) with tiny
    slices and groups, so that a small batch crosses many of them; shared targets,
    regions and a mixed (packed16 + int32 + direct) batch.  Same answers as with one
    slice, and as the oracle."""
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    opt = Optimal(eng, model, scoring)
    rng = random.Random(99)
    qs, ts = [], []
    shared_t = helpers.rand_dna(rng, 6000)
    for k in range(40):
        ql = rng.choice([20, 150, 400, 1000, 1300])       # 1300 > one sweep: int32 kernel
        tl = rng.choice([300, 2500, 9000])
        q, t = helpers.dna_pair(88000 + k, ql, tl)
        qs.append(q)
        ts.append(shared_t if k % 7 == 0 else t)
    pairs = PairSet(qs, ts)
    want_scores, want_paths = opt.find_score(pairs), opt.find_path(pairs)
    monkeypatch.setenv("C4B_STAGE_SLICE_KB", "8")
    monkeypatch.setenv("C4B_P1_MIN_GROUP", "2")
    got_scores, got_paths = opt.find_score(pairs), opt.find_path(pairs)
    monkeypatch.delenv("C4B_STAGE_SLICE_KB")
    monkeypatch.delenv("C4B_P1_MIN_GROUP")
    assert got_scores == want_scores and got_paths == want_paths
    for k in range(0, pairs.n, 3):
        want = oracle_path(model, scoring, qs[k], ts[k])
        assert got_scores[k] == want["score"], k
        assert got_paths[k]["region"] == want["region"] and got_paths[k]["ops"] == want["ops"], k


def test_long_queries_pipelined_sweeps(eng, params, scoring, monkeypatch):
    """Queries of many 1024-row sweeps: the warps of a CTA run the sweeps of one lattice
    (int32 kernel) or of one pair of lattices (packed score pass, queries up to 6399 bp)
    concurrently, hand-off row published through a shared-memory counter.  Same answers
    with 1, 3 and 8 warps, with the packed kernels off, and as the oracle; local and
    global scopes; ragged batch (partners with different sweep counts)."""
    from exonerate_b200 import Optimal, PairSet
    for name in ("affine_local_dna", "affine_global_dna"):
        model, _ = helpers.load_model(name, params)
        opt = Optimal(eng, model, scoring)
        qs, ts = [], []
        for k, (ql, tl) in enumerate([(5000, 3000), (9000, 1200), (1025, 4000), (3000, 3000), (40, 500), (2049, 2049)]):
            q, t = helpers.dna_pair(61000 + k, ql, tl, rate=0.1)
            qs.append(q)
            ts.append(t)
        pairs = PairSet(qs, ts)
        got = {}
        for w in ("1", "3", "8"):
            monkeypatch.setenv("C4B_AFFINE_WARPS", w)
            got[w] = (opt.find_score(pairs), opt.find_path(pairs))
        monkeypatch.delenv("C4B_AFFINE_WARPS")
        monkeypatch.setenv("C4B_AFFINE_PACK16", "0")   # int32 kernels only
        got["int32"] = (opt.find_score(pairs), opt.find_path(pairs))
        monkeypatch.delenv("C4B_AFFINE_PACK16")
        assert got["1"] == got["3"] == got["8"] == got["int32"]
        for k in range(pairs.n):
            want = oracle_path(model, scoring, qs[k], ts[k])
            assert got["8"][0][k] == want["score"], (name, k)
            assert got["8"][1][k]["region"] == want["region"] and got["8"][1][k]["ops"] == want["ops"], (name, k)
    # the packed multi-sweep score pass keeps the sweep count and takes the fewest rows per lane that cover
    # the longest query: 20 / 24 / 28 rows per lane (1201 / 2050 / 2601 lattice rows)
    from exonerate_b200 import Batch
    model, _ = helpers.load_model("affine_local_dna", params)
    opt = Optimal(eng, model, scoring)
    for maxq, rows in ((1200, 20), (2049, 24), (2600, 28)):
        qs, ts = [], []
        for k, (ql, tl) in enumerate([(maxq, 1500), (maxq // 2, 900), (maxq - 7, 2100)]):
            q, t = helpers.dna_pair(62000 + maxq + k, ql, tl, rate=0.1)
            qs.append(q)
            ts.append(t)
        pairs = PairSet(qs, ts)
        b = Batch(eng, model, scoring, pairs, want_path=False)
        assert "%d rows/lane" % rows in b.description, b.description
        b.close()
        scores, paths = opt.find_score(pairs), opt.find_path(pairs)
        for k in range(pairs.n):
            want = oracle_path(model, scoring, qs[k], ts[k])
            assert scores[k] == want["score"] and paths[k]["region"] == want["region"] and paths[k]["ops"] == want["ops"], (maxq, k)


def test_mixed_alphabets_share_a_batch(eng, params, scoring):
    """Queries of A/C/G/T only take the packed kernels (PRMT classes 0..3), queries with N
    or IUPAC codes the int32 kernels, in the same batch and launch sequence; long-target
    (two-pass) and short-target (single pass) routes."""
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    opt = Optimal(eng, model, scoring)
    rng = random.Random(123)
    qs, ts = [], []
    for k in range(14):
        ql = rng.choice([60, 300, 900])
        tl = rng.choice([400, 1200, 9000])
        q, t = helpers.dna_pair(33000 + k, ql, tl, rate=0.1)
        if k % 2:   # sprinkle ambiguity codes into every other query
            q = "".join(c if rng.random() > 0.03 else rng.choice("NRYK") for c in q)
        qs.append(q)
        ts.append(t)
    pairs = PairSet(qs, ts)
    scores, paths = opt.find_score(pairs), opt.find_path(pairs)
    for k in range(pairs.n):
        want = oracle_path(model, scoring, qs[k], ts[k])
        assert scores[k] == want["score"], k
        assert paths[k]["score"] == want["score"] and paths[k]["region"] == want["region"], k
        assert paths[k]["ops"] == want["ops"], k


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_packed_kernels_fuzz_penalties_and_matrices(eng, params, scoring, monkeypatch, seed):
    """Differential fuzz of the exactness guards of the packed kernels: random gap
    penalties (incl. open == extend, the -24 limit of the tagged traceback pass and beyond
    it), random match / mismatch / N scores (incl. mismatch below gap open, which the
    offset-binary variant must refuse), ragged shapes.  Packed path == int32 path, and a
    sample == oracle."""
    import copy
    from exonerate_b200 import Optimal, PairSet
    from exonerate_b200.models import default_params, host_model
    rng = random.Random(900 + seed)
    for trial in range(4):
        hp = default_params()
        hp.gap_extend = -rng.choice([1, 2, 4, 9, 24, 30])
        hp.gap_open = hp.gap_extend - rng.choice([0, 1, 8, 20, 60])
        model, _ = host_model("affine:local", params=hp)
        sc = copy.deepcopy(scoring)
        match, mism, nsc = rng.choice([1, 2, 5, 9]), -rng.choice([1, 3, 4, 15, 40]), rng.choice([0, -1, -2])
        for a in "ACGTN":
            for b_ in "ACGTN":
                v = nsc if "N" in (a, b_) else (match if a == b_ else mism)
                sc.dna_matrix[sc.dna_index[ord(a)] * 24 + sc.dna_index[ord(b_)]] = v
        qs, ts = [], []
        for k in range(11):
            ql, tl = rng.choice([(30, 200), (250, 300), (700, 900), (1000, 6000), (1023, 1024), (90, 9000)])
            q, t = helpers.dna_pair(seed * 1000 + trial * 50 + k, ql, tl, rate=rng.choice([0.0, 0.08, 0.25]))
            if k == 3:
                t = t[:len(t) // 2] + "N" * 7 + t[len(t) // 2:]
            qs.append(q)
            ts.append(t)
        pairs = PairSet(qs, ts)
        opt = Optimal(eng, model, sc)
        got = (opt.find_score(pairs), opt.find_path(pairs))
        monkeypatch.setenv("C4B_AFFINE_PACK16", "0")
        want = (opt.find_score(pairs), opt.find_path(pairs))
        monkeypatch.delenv("C4B_AFFINE_PACK16")
        assert got == want, (seed, trial, hp.gap_open, hp.gap_extend, match, mism, nsc)
        for k in (0, 1, 2):
            ref = oracle_path(model, sc, qs[k], ts[k])
            assert got[0][k] == ref["score"] and got[1][k]["ops"] == ref["ops"], (seed, trial, k)


def test_degenerate_lattices(eng, params, scoring):
    """Empty and one-symbol regions (query_length or target_length 0 / 1) mixed with normal
    lattices, on the packed affine kernels and the packed est2genome kernel."""
    from exonerate_b200 import Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    q, t = helpers.dna_pair(4711, 120, 400)
    regions = [(0, 0, 0, 400), (0, 0, 120, 0), (5, 7, 0, 0), (0, 0, 1, 1), (3, 10, 1, 300), (0, 0, 120, 1),
               (0, 0, 120, 400), (119, 399, 1, 1)]
    for name in ("affine_local_dna", "affine_global_dna", "est2genome"):
        model, _ = helpers.load_model(name, params)
        sp = splice_arrays(t) if name == "est2genome" else None
        pairs = PairSet([q] * len(regions), [t] * len(regions), splice=[sp] * len(regions), regions=regions)
        opt = Optimal(eng, model, scoring)
        scores, paths = opt.find_score(pairs), opt.find_path(pairs)
        for k, reg in enumerate(regions):
            want = helpers.oracle_viterbi(model, scoring, helpers.PairBuf(q, t, splice=sp, region=reg),
                                          abi.MODE_FIND_PATH)
            assert scores[k] == want["score"], (name, reg)
            assert paths[k]["score"] == want["score"] and paths[k]["region"] == want["region"], (name, reg)
            assert paths[k]["ops"] == want["ops"], (name, reg)


def test_protein_smem_scoring_vs_oracle(eng, params, scoring):
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_protein", params)
    qs, ts = [], []
    for k, (ql, tl) in enumerate([(30, 60), (200, 400), (500, 500), (700, 1500)]):
        q, t = helpers.protein_pair(600 + k, ql, tl)
        qs.append(q)
        ts.append(t)
    pairs = PairSet(qs, ts)
    opt = Optimal(eng, model, scoring)
    paths = opt.find_path(pairs)
    for k in range(pairs.n):
        want = oracle_path(model, scoring, qs[k], ts[k])
        assert paths[k]["score"] == want["score"] and paths[k]["region"] == want["region"]
        assert paths[k]["ops"] == want["ops"]


def test_regions_sub_lattices(eng, params, scoring):
    """A c4b_pair may address a Region of the sequences (src/c4/region.h)."""
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    q, t = helpers.dna_pair(42, 300, 900)
    regions = [(0, 0, 300, 900), (10, 100, 200, 500), (150, 0, 150, 900), (0, 450, 300, 450)]
    pairs = PairSet([q] * 4, [t] * 4, regions=regions)
    opt = Optimal(eng, model, scoring)
    paths = opt.find_path(pairs)
    for k, reg in enumerate(regions):
        want = helpers.oracle_viterbi(model, scoring, helpers.PairBuf(q, t, region=reg), abi.MODE_FIND_PATH)
        assert paths[k]["score"] == want["score"] and paths[k]["region"] == want["region"]
        assert paths[k]["ops"] == want["ops"]


def blocked_points(model, alignments):
    pts, seen = [], set()
    for a in alignments:
        qp, tp = a["region"][0], a["region"][1]
        for tid, length in a["ops"]:
            tr = model.transitions[tid]
            for _ in range(length):
                if tr.label == abi.LABEL_MATCH and (qp, tp) not in seen:
                    seen.add((qp, tp))
                    pts.append((qp, tp))
                qp += tr.advance_query
                tp += tr.advance_target
    return pts


def test_subopt_blocking(eng, params, scoring):
    """SubOpt_Index blocked cells (generic path): the reference's own series."""
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    opt = Optimal(eng, model, scoring)
    n = 0
    for case in helpers.load_cases("affine_local_dna"):
        series = case.get("subopt_series")
        if not series:
            continue
        done = []
        for ref in series:
            pairs = PairSet([case["q"]], [case["t"]], blocked=[blocked_points(model, done)])
            r = opt.find_path(pairs)[0]
            assert r["score"] == ref["score"] and r["region"] == ref["region"]
            assert r["ops"] == [tuple(o) for o in ref["ops"]]
            done.append(ref)
            n += 1
    assert n >= 6


def _subopt_series(opt, model, scoring, q, t, rounds, region=None):
    """the --subopt loop of GAM_Result_exhaustive_create (src/hub/gam.c:1160-1172) against the
    oracle: every iteration blocks the match cells of all earlier paths"""
    from exonerate_b200 import PairSet
    done = []
    for _ in range(rounds):
        pts = blocked_points(model, done)
        if region:   # blocked lists are in REGION coordinates (src/c4/subopt.c:250-338)
            pts = [(a - region[0], b - region[1]) for a, b in pts
                   if 0 <= a - region[0] <= region[2] and 0 <= b - region[1] <= region[3]]
        want = helpers.oracle_viterbi(model, scoring, helpers.PairBuf(q, t, blocked=pts, region=region),
                                      abi.MODE_FIND_PATH, max_ops=len(q) + len(t) + 8)
        got = opt.find_path(PairSet([q], [t], blocked=[pts], regions=[region] if region else None))[0]
        assert got["score"] == want["score"] and got["region"] == want["region"], (got["score"], want["score"])
        assert got["ops"] == want["ops"]
        if not want["ops"]:
            break
        done.append(want)
    return done


@pytest.mark.parametrize("route", ["box", "direct", "windows"])
def test_subopt_on_the_systolic_table_driven_kernel(eng, params, scoring, monkeypatch, route):
    """SubOpt blocked cells on the systolic specialisation (JIT_SYS_BLK: the callers' lists as {column, row
    mask} entries per lane strip, rebuilt for the alignment boxes): the --subopt series against the oracle for
    protein2genome, coding2coding, est2genome and affine:local forced onto the table-driven path -- REGION +
    box, one PATH pass over the full lattice, and column windows; several strips, a sub-region with its own
    list.  (CLI default --subopt yes: every iteration after the first carries such a list.)"""
    from exonerate_b200 import Batch, Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    monkeypatch.setenv("C4B_FORCE_GENERIC", "1")
    monkeypatch.setenv("C4B_GENERIC_JIT", "1")
    if route == "windows":
        monkeypatch.setenv("C4B_GENERIC_TB_BUDGET_KB", "1")
        monkeypatch.setenv("C4B_GENERIC_WINDOW_COLS", "64")
    else:
        monkeypatch.setenv("C4B_GENERIC_DIRECT_PATH", "1" if route == "direct" else "0")
    for name in ("protein2genome", "coding2coding", "est2genome", "affine_local_dna"):
        model, _ = helpers.load_model(name, params)
        if name == "protein2genome":
            q, t = bench_p2g_pair(8100, 150, 3000)
        elif name == "est2genome":
            q, t = helpers.gene_pair(8101, 400, 4000, n_exons=3, rate=0.02)
        else:
            q, t = helpers.dna_pair(8102, 420 if name == "coding2coding" else 700, 2500)
        sp = splice_arrays(t) if name in ("protein2genome", "est2genome") else None

        class Opt:   # _subopt_series builds its own PairSet: give it the splice arrays
            def find_path(self, pairs):
                ps = PairSet([q], [t], splice=[sp] if sp is not None else None, blocked=[self.pts], regions=self.regions)
                if len(self.pts) and not self.checked:
                    b = Batch(eng, model, scoring, ps, want_path=True)
                    b.run()
                    assert b.kernel_name == "generic_jit_systolic", name
                    b.close()
                    self.checked = True
                return Optimal(eng, model, scoring).find_path(ps)
        o = Opt()
        o.checked = False
        done = []
        for region in (None, (len(q) // 8, len(t) // 10, len(q) - len(q) // 8 - 3, len(t) - len(t) // 10 - 7)):
            done = []
            for _ in range(3):
                pts = blocked_points(model, done)
                if region:
                    pts = [(a - region[0], b - region[1]) for a, b in pts
                           if 0 <= a - region[0] <= region[2] and 0 <= b - region[1] <= region[3]]
                o.pts, o.regions = pts, [region] if region else None
                want = helpers.oracle_viterbi(model, scoring, helpers.PairBuf(q, t, splice=sp, blocked=pts, region=region),
                                              abi.MODE_FIND_PATH, max_ops=len(q) + len(t) + 8)
                got = o.find_path(None)[0]
                assert got["score"] == want["score"] and got["region"] == want["region"], (name, route, region)
                assert got["ops"] == want["ops"], (name, route, region)
                if not want["ops"]:
                    break
                done.append(want)
        assert o.checked and len(done) >= 1, name


@pytest.mark.parametrize("name,shape", [("affine_local_dna", (300, 900)), ("affine_local_dna", (1000, 20000)),
                                        ("affine_local_dna", (1500, 1700)), ("affine_global_dna", (120, 150)),
                                        ("affine_bestfit_dna", (90, 400)), ("affine_overlap_dna", (200, 260)),
                                        ("affine_local_protein", (250, 600))])
def test_subopt_on_the_affine_kernels(eng, params, scoring, name, shape):
    """SubOpt blocked cells stay on the affine kernels (per-lattice routing; the int32 kernel's
    BLK variant masks T4 at blocked destination cells, viterbi.c:701-704): four iterations of the
    sub-optimal series, single pass and banded two-pass, one and several sweeps, all scopes."""
    from exonerate_b200 import Batch, Optimal, PairSet
    model, _ = helpers.load_model(name, params)
    opt = Optimal(eng, model, scoring)
    if name.endswith("protein"):
        q, t = helpers.protein_pair(7100, *shape)
    else:
        q, t = helpers.dna_pair(7100 + shape[0], *shape)
    done = _subopt_series(opt, model, scoring, q, t, 4)
    assert len(done) >= 2
    b = Batch(eng, model, scoring, PairSet([q], [t], blocked=[blocked_points(model, done[:1])]))
    assert b.kernel_name == "affine_systolic"
    b.close()


def test_subopt_mixed_batch_and_regions(eng, params, scoring):
    """blocked and unblocked lattices in ONE batch: the unblocked ones keep the packed 16-bit
    kernels, the blocked ones take the int32 BLK kernel; a sub-region with its own list"""
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    opt = Optimal(eng, model, scoring)
    qs, ts, blocked, want = [], [], [], []
    for k in range(24):
        q, t = helpers.dna_pair(7300 + k, 200 + 37 * k, 3000 + 811 * k)
        first = helpers.oracle_viterbi(model, scoring, helpers.PairBuf(q, t), abi.MODE_FIND_PATH,
                                       max_ops=len(q) + len(t) + 8)
        pts = blocked_points(model, [first]) if k % 3 else []
        qs.append(q)
        ts.append(t)
        blocked.append(pts)
        want.append(helpers.oracle_viterbi(model, scoring, helpers.PairBuf(q, t, blocked=pts), abi.MODE_FIND_PATH,
                                           max_ops=len(q) + len(t) + 8))
    got = opt.find_path(PairSet(qs, ts, blocked=blocked))
    scores = opt.find_score(PairSet(qs, ts, blocked=blocked))
    for k in range(24):
        assert got[k]["score"] == want[k]["score"] == scores[k], k
        assert got[k]["region"] == want[k]["region"] and got[k]["ops"] == want[k]["ops"], k
    q, t = helpers.dna_pair(7400, 400, 5000)
    _subopt_series(opt, model, scoring, q, t, 3, region=(20, 300, 350, 4500))


def test_pinned_caller_buffers_skip_the_bounce(eng, params, scoring):
    """C4B_PAIR_BUFFERS_PINNED: sequences in page-locked caller memory are DMA'd directly -- rows of
    one array as 2-D copies, ragged buffers one by one -- with the same results as the bounce path."""
    import torch
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    opt = Optimal(eng, model, scoring)
    n, ql, tl = 40, 300, 5000
    tq = torch.empty((n, ql), dtype=torch.uint8, pin_memory=True)
    tt = torch.empty((n, tl), dtype=torch.uint8, pin_memory=True)
    for k in range(n):
        q, t = helpers.dna_pair(9500 + k, ql, tl)
        tq.numpy()[k] = np.frombuffer(q.encode(), dtype=np.uint8)
        tt.numpy()[k] = np.frombuffer(t.encode(), dtype=np.uint8)
    rows_q, rows_t = [tq.numpy()[k] for k in range(n)], [tt.numpy()[k] for k in range(n)]
    want = opt.find_path(PairSet(rows_q, rows_t))
    assert opt.find_path(PairSet(rows_q, rows_t, pinned=True)) == want
    # ragged: views of different lengths into the same pinned arrays, shuffled order
    order = list(range(n))
    random.Random(3).shuffle(order)
    rq = [tq.numpy()[k][: 100 + 5 * k] for k in order]
    rt = [tt.numpy()[k][: 2000 + 70 * k] for k in order]
    assert opt.find_path(PairSet(rq, rt, pinned=True)) == opt.find_path(PairSet(rq, rt))
    assert opt.find_score(PairSet(rq, rt, pinned=True)) == [w["score"] for w in opt.find_path(PairSet(rq, rt))]


def test_device_group_shards_and_merges(eng, params, scoring):
    """c4b_group: the batch is dealt to the group's engines by cost, the shards run on one host
    thread per engine, results and op lists come back in pair order -- identical to one engine.
    (On a one-GPU box all three engines sit on device 0: the sharding / merging logic does not
    care which GPU; with more GPUs visible the group spans them.)"""
    import torch
    from exonerate_b200 import Group, Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    nd = max(1, torch.cuda.device_count())
    grp = Group([0, 1 % nd, 2 % nd])
    assert grp.size == 3
    for name in ("affine_local_dna", "est2genome", "coding2coding"):
        model, _ = helpers.load_model(name, params)
        qs, ts = [], []
        for k in range(17):
            if name == "est2genome":
                q, t = helpers.gene_pair(8100 + k, 100 + 40 * k, 2000 + 300 * k, n_exons=1 + k % 3)
            else:
                q, t = helpers.dna_pair(8100 + k, 90 + 31 * k, 700 + 517 * k)
            qs.append(q)
            ts.append(t)
        sp = [splice_arrays(t) for t in ts] if name == "est2genome" else None
        pairs = PairSet(qs, ts, splice=sp)
        want = Optimal(eng, model, scoring).find_path(pairs)
        got = grp.find_path(model, scoring, pairs)
        assert got == want, name
        assert grp.find_score(model, scoring, pairs) == [w["score"] for w in want], name
        assert grp.find_path(model, scoring, pairs, threshold=10 ** 6)[0]["status"] == 1
    assert grp.find_path(model, scoring, PairSet([], [])) == []
    assert grp.kernel_launches() > 0
    grp.close()


def test_threshold_and_errors(eng, params, scoring):
    from exonerate_b200 import C4BError, Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    opt = Optimal(eng, model, scoring)
    q, t = helpers.dna_pair(1, 50, 80)
    r = opt.find_path(PairSet([q], [t]), threshold=10 ** 6)[0]
    assert r["status"] == 1 and r["ops"] == []      # Optimal_find_path returns NULL
    with pytest.raises(C4BError):                   # symbol outside the matrix alphabet
        opt.find_path(PairSet(["ACGT-ACGT"], ["ACGTACGT"]))
    with pytest.raises(C4BError):                   # ops buffer too small is an error, not truncation
        opt.find_path(PairSet([q], [t]), ops_capacity=1)
    assert opt.find_path(PairSet([], [])) == []     # empty batch


@pytest.mark.parametrize("ql,tl", [(1000, 100000)])
def test_full_size_properties(eng, params, scoring, ql, tl):
    """BASELINE.json metric shape (1 kbp x 100 kbp): size-independent checks --
    the path re-scores to the DP score (Alignment_is_valid), stays inside the
    lattice, the score equals score-only mode; one lattice also against the
    oracle's full fill."""
    import ctypes as C
    from exonerate_b200 import Optimal, PairSet
    model, _ = helpers.load_model("affine_local_dna", params)
    qs, ts = [], []
    for k in range(6):
        q, t = helpers.dna_pair(31000 + k, ql, tl)
        qs.append(q)
        ts.append(t)
    pairs = PairSet(qs, ts)
    opt = Optimal(eng, model, scoring)
    scores = opt.find_score(pairs)
    paths = opt.find_path(pairs)
    lib = helpers.oracle()
    for k in range(pairs.n):
        r = paths[k]
        assert r["score"] == scores[k] > 2000
        qs_, ts_, qlen, tlen = r["region"]
        assert 0 <= qs_ and qs_ + qlen <= ql and 0 <= ts_ and ts_ + tlen <= tl
        res = abi.Result()
        res.score, res.query_start, res.target_start = r["score"], qs_, ts_
        res.n_ops = len(r["ops"])
        ops = np.array([x for op in r["ops"] for x in op], dtype=np.int32)
        pb = helpers.PairBuf(qs[k], ts[k])
        assert lib.c4o_rescore_path(C.byref(model), C.byref(scoring), C.byref(pb.pair), C.byref(res),
                                    ops.ctypes.data) == r["score"]
        aq = sum(model.transitions[t_].advance_query * n for t_, n in r["ops"])
        at = sum(model.transitions[t_].advance_target * n for t_, n in r["ops"])
        assert (aq, at) == (qlen, tlen)
    want = helpers.oracle_find_path(model, scoring, helpers.PairBuf(qs[0], ts[0]),
                                    region_threshold_cells=0, max_ops=ql + tl)
    assert paths[0]["score"] == want["score"] and paths[0]["region"] == want["region"]
    assert paths[0]["ops"] == want["ops"]


def e2g_oracle(model, scoring, q, t, sp, region=None):
    return helpers.oracle_viterbi(model, scoring, helpers.PairBuf(q, t, splice=sp, region=region),
                                  abi.MODE_FIND_PATH, max_ops=len(q) + len(t) + 8)


E2G_SHAPES = [(1, 1), (1, 50), (30, 2), (20, 120), (255, 1500), (256, 900), (257, 2500), (511, 1200),
              (513, 3000), (600, 5000), (1000, 6000), (1300, 2500), (2047, 2400)]


@pytest.mark.parametrize("kernel", ["e2g_packed16", "e2g_packed16:full", "e2g_packed16:rows8", "e2g_packed16:rows16",
                                    "e2g_packed16:rows8:full", "e2g_packed16:rows8:warps1", "e2g_packed16:wcols64",
                                    "e2g_packed16:rows4", "e2g_packed16:rows4:full", "e2g_packed16:rows4:warps3",
                                    "e2g_packed16:rows8:wcols1024", "e2g_systolic"])
def test_est2genome_systolic_vs_oracle(eng, params, scoring, monkeypatch, kernel):
    """The hand-specialised est2genome kernels -- e2g_packed16 (both strands per
    register, one warp per lattice, 512-row sweeps) and the int32 e2g_systolic
    (1..8 strips of 256 rows, pipelined through shared memory; C4B_E2G_PACK16=0)
    -- against the oracle: forward and reverse-strand genes,
    ragged sizes around the strip height, unrelated and N-rich inputs; one mixed
    batch (idle strips for short queries) and per-shape batches (every strip count)."""
    from exonerate_b200 import Batch, Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    model, _ = helpers.load_model("est2genome", params)
    rng = random.Random(3)
    qs, ts = [], []
    for k, (ql, tl) in enumerate(E2G_SHAPES):
        q, t = helpers.gene_pair(4000 + k, ql, tl, n_exons=1 + k % 5, rate=rng.choice([0.0, 0.02, 0.1]),
                                 reverse=bool(k & 1))
        qs.append(q)
        ts.append(t)
    qs.append(helpers.rand_dna(rng, 300)); ts.append(helpers.rand_dna(rng, 1500))
    qs.append(helpers.rand_dna(rng, 200, "ACGTN")); ts.append(helpers.rand_dna(rng, 800, "ACGTNRY"))
    sp = [splice_arrays(t) for t in ts]
    want = [e2g_oracle(model, scoring, q, t, s_) for q, t, s_ in zip(qs, ts, sp)]
    assert any(any(model.transitions[t_].advance_target == 2 for t_, _ in w["ops"]) for w in want)
    opt = Optimal(eng, model, scoring)
    if kernel == "e2g_systolic":
        monkeypatch.setenv("C4B_E2G_PACK16", "0")
    # rows per lane: 16 (512-row sweeps) or 8 (256-row sweeps on up to four pipelined warps, the
    # small-batch shape); unset = the library's choice by batch size
    for opt_ in kernel.split(":")[1:]:
        if opt_ == "full":           # records for the whole lattice instead of checkpoints + windows
            monkeypatch.setenv("C4B_E2G_WINDOWS", "0")
        elif opt_.startswith("rows"):
            monkeypatch.setenv("C4B_E2G_ROWS", opt_[4:])
        elif opt_.startswith("warps"):
            monkeypatch.setenv("C4B_E2G_WARPS", opt_[5:])
        elif opt_.startswith("wcols"):   # checkpoint window width of the windowed traceback (default: by memory)
            monkeypatch.setenv("C4B_E2G_WINDOW_COLS", opt_[5:])
    kernel = kernel.split(":")[0]

    def check(idx):
        pairs = PairSet([qs[k] for k in idx], [ts[k] for k in idx], splice=[sp[k] for k in idx])
        b = Batch(eng, model, scoring, pairs, want_path=True)
        assert b.kernel_name == kernel
        b.close()
        scores = opt.find_score(pairs)
        paths = opt.find_path(pairs)
        for n, k in enumerate(idx):
            shape = (len(qs[k]), len(ts[k]))
            assert scores[n] == want[k]["score"], shape
            assert paths[n]["score"] == want[k]["score"], shape
            assert paths[n]["region"] == want[k]["region"], shape
            assert paths[n]["ops"] == want[k]["ops"], shape

    check(list(range(len(qs))))
    for k in range(len(E2G_SHAPES)):
        check([k])


def test_est2genome_intron_gain_leaves_the_packed_kernel(eng, params, scoring):
    """--intronpenalty 0 (accepted by the reference, intron.c:35): an intron then GAINS score at good
    splice sites, introns chain without consuming query and values are no longer bounded by the
    16-bit argument -- the batch must leave the packed kernel and still be exact."""
    from exonerate_b200 import Batch, Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    base, _ = helpers.load_model("est2genome", params)
    model = type(base).from_buffer_copy(base)
    n_pre = 0
    for k in range(model.n_calcs):
        if model.calcs[k].kind == abi.CALC_SPLICE_PRE:
            model.calcs[k].param[0] = 0
            n_pre += 1
    assert n_pre >= 2
    qs, ts = [], []
    for k, (ql, tl) in enumerate([(120, 1500), (300, 2500), (700, 4000)]):
        q, t = helpers.gene_pair(9100 + k, ql, tl, n_exons=3, rate=0.02, reverse=bool(k & 1))
        qs.append(q)
        ts.append(t)
    sp = [splice_arrays(t) for t in ts]
    pairs = PairSet(qs, ts, splice=sp)
    b = Batch(eng, model, scoring, pairs, want_path=True)
    assert b.kernel_name != "e2g_packed16"
    b.close()
    opt = Optimal(eng, model, scoring)
    scores, paths = opt.find_score(pairs), opt.find_path(pairs)
    for k in range(pairs.n):
        want = e2g_oracle(model, scoring, qs[k], ts[k], sp[k])
        assert scores[k] == want["score"] and paths[k]["score"] == want["score"], k
        assert paths[k]["region"] == want["region"] and paths[k]["ops"] == want["ops"], k
    # the default penalty keeps the packed kernel
    b = Batch(eng, base, scoring, pairs, want_path=True)
    assert b.kernel_name == "e2g_packed16"
    b.close()


@pytest.mark.parametrize("rows", ["4", "8", "16"])
def test_est2genome_windowed_traceback_long_introns(eng, params, scoring, monkeypatch, rows):
    """(rows per lane 8: the small-batch shape, window refills on independent warps.)
    find_path on long targets: pass 1 saves column checkpoints every 1024 columns, the
    traceback refills only the windows under the path and crosses introns in one jump
    (the checkpoint holds the intron's age).  Introns of 3 .. 40 kbp (the 16-bit age
    saturates at 32767: no jump, window-by-window walk), both strands; against the
    full-lattice record pass and, for the smaller ones, the oracle."""
    from exonerate_b200 import Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    model, _ = helpers.load_model("est2genome", params)
    opt = Optimal(eng, model, scoring)
    qs, ts = [], []
    for k, (ql, tl, nex) in enumerate([(300, 12000, 3), (600, 30000, 4), (1000, 100000, 5), (900, 70000, 2),
                                       (200, 90000, 2), (1000, 100000, 5), (520, 45000, 3)]):
        q, t = helpers.gene_pair(6100 + k, ql, tl, n_exons=nex, rate=0.02, reverse=bool(k & 1))
        qs.append(q)
        ts.append(t)
    sp = [splice_arrays(t) for t in ts]
    pairs = PairSet(qs, ts, splice=sp)
    monkeypatch.setenv("C4B_E2G_ROWS", rows)
    got = opt.find_path(pairs)
    monkeypatch.setenv("C4B_E2G_WINDOWS", "0")
    want = opt.find_path(pairs)
    monkeypatch.delenv("C4B_E2G_WINDOWS")
    assert got == want
    assert all(any(model.transitions[t_].advance_target == 2 for t_, _ in r["ops"]) for r in got)
    for k in (0, 1):
        ref = e2g_oracle(model, scoring, qs[k], ts[k], sp[k])
        assert got[k]["score"] == ref["score"] and got[k]["region"] == ref["region"] and got[k]["ops"] == ref["ops"]


def test_est2genome_regions_threshold_and_fallback(eng, params, scoring, monkeypatch):
    from exonerate_b200 import Batch, Optimal, PairSet
    from exonerate_b200.models import splice_arrays
    model, _ = helpers.load_model("est2genome", params)
    opt = Optimal(eng, model, scoring)
    q, t = helpers.gene_pair(77, 400, 3000)
    sp = splice_arrays(t)
    regions = [(0, 0, 400, 3000), (10, 100, 300, 2500), (200, 0, 200, 3000), (0, 1500, 400, 1500)]
    paths = opt.find_path(PairSet([q] * 4, [t] * 4, splice=[sp] * 4, regions=regions))
    for k, reg in enumerate(regions):
        want = e2g_oracle(model, scoring, q, t, sp, region=reg)
        assert paths[k]["score"] == want["score"] and paths[k]["region"] == want["region"], reg
        assert paths[k]["ops"] == want["ops"], reg
    r = opt.find_path(PairSet([q], [t], splice=[sp]), threshold=10 ** 6)[0]
    assert r["status"] == 1 and r["ops"] == []
    # a query of five 512-row sweeps on the packed kernel; without it, queries beyond 8
    # strips take the table-driven kernel: same answers
    q2, t2 = helpers.gene_pair(78, 2100, 2600)
    sp2 = splice_arrays(t2)
    pairs = PairSet([q2], [t2], splice=[sp2])
    want = e2g_oracle(model, scoring, q2, t2, sp2)
    for env, name in ((None, "e2g_packed16"), ("0", "generic_wavefront")):
        if env is not None:
            monkeypatch.setenv("C4B_E2G_PACK16", env)
        b = Batch(eng, model, scoring, pairs, want_path=True)
        assert b.kernel_name == name
        b.close()
        got = opt.find_path(pairs)[0]
        assert got["score"] == want["score"] and got["ops"] == want["ops"], name
        assert opt.find_score(pairs)[0] == want["score"], name


# ---------------------------------------------------------------------------
# HSP seeding / extension
# ---------------------------------------------------------------------------
def test_hsp_extend_golden_and_oracle(eng, scoring):
    """c4b_hsp_extend_batch + the HSPset binding against the reference's own HSP lists
    (tests/golden/hsp_cases.json: DNA, protein, protein vs translated DNA, soft-masked)
    and, seed by seed, against the oracle."""
    import json
    from exonerate_b200 import HSPset
    cases = json.load(open(helpers.GOLDEN + "/hsp_cases.json"))
    total = 0
    for case in cases:
        param = helpers.hsp_param(case)
        qm = helpers.softmask_bytes(case["q"], case["softmask_query"])
        tm = helpers.softmask_bytes(case["t"], case["softmask_target"])
        hs = HSPset(eng, scoring, param, case["q"], case["t"], qm, tm)
        for qs, ts in case["seeds"]:
            hs.seed_hsp(qs, ts)
        ext = hs.extend_all()
        want = helpers.oracle_hsp_extend(scoring, param, case["q"], case["t"], [tuple(s) for s in case["seeds"]], qm, tm)
        for k in range(len(case["seeds"])):
            for f, _ in abi.Hsp._fields_:
                assert getattr(ext[k], f) == getattr(want[k], f), (case["name"], k, f)
        assert hs.finalise() == case["hsps"], case["name"]
        total += len(case["hsps"])
    assert total >= 150


def test_hsp_extend_large_random_vs_oracle(eng, scoring):
    """100k seeds on a 20 kbp x 200 kbp comparison (every 12-mer match + noise), incl.
    seeds whose trimmed score is negative (status 1: fatal in the reference) and masks."""
    from exonerate_b200 import C4BError, HSPset
    rng = random.Random(4242)
    q, t = helpers.dna_pair(99001, 20000, 200000, rate=0.12)
    t = t[:50000] + t[50000:50400].lower() + t[50400:]
    param = abi.HspParam(abi.CALC_MATCH_DNA, 12, 30, 75)
    tm = helpers.softmask_bytes(t, True)
    index = {}
    for i in range(len(q) - 11):
        index.setdefault(q[i:i + 12], []).append(i)
    seeds = []
    for j in range(len(t) - 11):
        for i in index.get(t[j:j + 12].upper(), ()):
            seeds.append((i, j))
    noise = [(rng.randrange(0, len(q) - 12), rng.randrange(0, len(t) - 12)) for _ in range(3000)]
    hs = HSPset(eng, scoring, param, q, t, None, tm)
    for s_ in seeds + noise:
        hs.seed_hsp(*s_)
    ext = hs.extend_all()
    pick = list(range(0, len(seeds), max(1, len(seeds) // 400))) + list(range(len(seeds), len(seeds) + len(noise), 7))
    want = helpers.oracle_hsp_extend(scoring, param, q, t, [(seeds + noise)[k] for k in pick], None, tm)
    for n, k in enumerate(pick):
        for f, _ in abi.Hsp._fields_:
            assert getattr(ext[k], f) == getattr(want[n], f), (k, f)
    assert any(ext[k].status == 1 for k in range(len(seeds), len(seeds) + len(noise)))
    with pytest.raises(C4BError):
        hs.finalise()
    hs2 = HSPset(eng, scoring, param, q, t, None, tm)
    for s_ in seeds:
        hs2.seed_hsp(*s_)
    got = hs2.finalise()
    ext2 = helpers.oracle_hsp_extend(scoring, param, q, t, seeds, None, tm)
    assert got == helpers.hspset_replay(param, len(q), seeds, ext2) and len(got) >= 1
    with pytest.raises(C4BError):   # a seed outside the sequences / a foreign symbol are errors
        h3 = HSPset(eng, scoring, param, q, t)
        h3.seed_hsp(len(q) - 3, 0)
        h3.finalise()
    with pytest.raises(C4BError):
        h4 = HSPset(eng, scoring, param, "ACGT-ACGTACGTACG", "ACGTACGTACGTACGT")
        h4.seed_hsp(0, 0)
        h4.finalise()


# ---------------------------------------------------------------------------
# BSDP derived models: cell callbacks as tables
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["est2genome", "affine_local_dna", "protein2genome"])
def test_cell_callback_tables_vs_oracle(eng, params, scoring, name):
    """c4b_viterbi_calculate_cells: START's score and shadow slots per cell come from a
    table (what cell_start_func returns), END's cell is returned wherever END is reached
    (what cell_end_func is handed) -- random tables incl. impossible starts, FIND_SCORE
    and FIND_PATH, against the oracle."""
    from exonerate_b200 import PairSet
    from exonerate_b200.engine import viterbi_calculate_cells
    from exonerate_b200.models import splice_arrays
    model, _ = helpers.load_model(name, params)
    rng = np.random.default_rng(17)
    C_ = 1 + model.n_shadow_slots
    for trial, (ql, tl) in enumerate([(40, 90), (25, 300), (70, 70)]):
        if name == "protein2genome":
            q, _t = helpers.protein_pair(5100 + trial, ql, 30)
            t = helpers.rand_dna(random.Random(trial), tl)
        else:
            q, t = helpers.gene_pair(5000 + trial, ql, tl, n_exons=2) if name == "est2genome" else \
                helpers.dna_pair(5000 + trial, ql, tl)
        sp = splice_arrays(t) if name in ("est2genome", "protein2genome") else None
        cells = (len(q) + 1) * (len(t) + 1)
        start = np.zeros((cells, C_), dtype=np.int32)
        start[:, 0] = rng.integers(-40, 60, cells)
        start[rng.random(cells) < 0.5, 0] = abi.IMPOSSIBLY_LOW_SCORE
        if C_ > 1:   # shadow slots are sequence positions (<= the cell's own column)
            cols = np.tile(np.arange(len(t) + 1), len(q) + 1)
            for l in range(1, C_):
                start[:, l] = (cols * rng.random(cells)).astype(np.int32)
        pairs = PairSet([q], [t], splice=[sp])
        pb = helpers.PairBuf(q, t, splice=sp)
        for mode in (abi.MODE_FIND_SCORE, abi.MODE_FIND_PATH):
            end_got = np.full((cells, C_), -7, dtype=np.int32)
            end_want = np.full((cells, C_), -7, dtype=np.int32)
            got = viterbi_calculate_cells(eng, model, scoring, pairs, mode, start, end_got)
            want = helpers.oracle_viterbi_cells(model, scoring, pb, mode, start, end_want)
            assert got["score"] == want["score"], (name, trial, mode)
            assert (end_got == end_want).all(), (name, trial, mode)
            if mode == abi.MODE_FIND_PATH:
                assert got["region"] == want["region"] and got["ops"] == want["ops"], (name, trial)
        # without tables the call is the plain fill
        plain = viterbi_calculate_cells(eng, model, scoring, pairs, abi.MODE_FIND_PATH)
        ref = helpers.oracle_viterbi(model, scoring, pb, abi.MODE_FIND_PATH)
        assert plain["score"] == ref["score"] and plain["ops"] == ref["ops"]


@pytest.mark.parametrize("name", ["est2genome", "affine_local_dna", "protein2genome"])
def test_span_score_batch_vs_oracle(eng, params, scoring, name, monkeypatch):
    """c4b_span_score_batch = SAR_Span_find_score for many span edges at once: src fill with END's
    cell reported from every cell, Heuristic_Span_integrate, the START table of
    Heuristic_Span_dst_init_start_func, dst fill -- against the same pipeline on the oracle
    (c4o_viterbi_cells + c4o_span_integrate), interpreter and specialised kernels."""
    from exonerate_b200 import PairSet
    import test_span_oracle
    from exonerate_b200.engine import span_score_batch
    from exonerate_b200.models import splice_arrays
    model, _ = helpers.load_model(name, params)
    C_ = 1 + model.n_shadow_slots
    rng = random.Random(41)
    qs, ts, sp, sreg, dreg, spans, want = [], [], [], [], [], [], []
    for k in range(14):
        ql, tl = rng.choice([60, 90, 140]), rng.choice([400, 700, 1200])
        if name == "protein2genome":
            q, _t = helpers.protein_pair(6100 + k, ql, 30)
            t = helpers.rand_dna(random.Random(k), tl)
        else:
            q, t = helpers.gene_pair(6000 + k, ql, tl, n_exons=2) if name == "est2genome" else helpers.dna_pair(6000 + k, ql, tl)
        s_ = splice_arrays(t) if name in ("est2genome", "protein2genome") else None
        # a src box near the start, a dst box further along the target, overlapping in the query
        sq, st = rng.randrange(0, 10), rng.randrange(0, 40)
        sl, stl = rng.randrange(12, 30), rng.randrange(20, 70)
        dq, dt = sq + rng.randrange(0, sl), st + stl + rng.randrange(0, 200)
        dl, dtl = rng.randrange(10, min(28, len(q) - dq)), rng.randrange(20, min(70, len(t) - dt))
        span = (0, rng.choice([0, 0, 5]), rng.choice([0, 20]), rng.choice([100, 200000]))
        qs.append(q); ts.append(t); sp.append(s_)
        sreg.append((sq, st, sl, stl)); dreg.append((dq, dt, dl, dtl)); spans.append(span)
        # the oracle pipeline
        src_end = np.full(((sl + 1) * (stl + 1), C_), 0, dtype=np.int32)
        src_end[:, 0] = abi.IMPOSSIBLY_LOW_SCORE     # Heuristic_Span_clear
        helpers.oracle_viterbi_cells(model, scoring, helpers.PairBuf(q, t, splice=s_, region=sreg[-1]),
                                     abi.MODE_FIND_SCORE, None, src_end)
        pos = test_span_oracle.oracle_span_integrate({"src_scores": src_end[:, 0].tolist(), "src_region": sreg[-1],
                                                      "dst_region": dreg[-1], "span": span})
        start = np.zeros(((dl + 1) * (dtl + 1), C_), dtype=np.int32)
        start[:, 0] = abi.IMPOSSIBLY_LOW_SCORE
        for c in range((dl + 1) * (dtl + 1)):
            if pos[2 * c] >= 0:
                start[c] = src_end[(pos[2 * c] - sq) * (stl + 1) + (pos[2 * c + 1] - st)]
        w = helpers.oracle_viterbi_cells(model, scoring, helpers.PairBuf(q, t, splice=s_, region=dreg[-1]),
                                         abi.MODE_FIND_SCORE, start, None)
        want.append(w["score"])
    pairs = PairSet(qs, ts, splice=sp)
    for jit in ("0", "1"):
        monkeypatch.setenv("C4B_GENERIC_JIT", jit)
        got = span_score_batch(eng, model, model, scoring, pairs, sreg, dreg, spans)
        assert got == want, (name, jit)
    assert len(set(want)) > 3


def test_span_integrate_golden(eng):
    """c4b_span_integrate against the vectors the unmodified Heuristic_Span_integrate produced
    (tests/golden/make_span_golden.py): every dst cell's source position, ties and empty
    windows included -- SURVEY 8f row 1."""
    import json
    import os
    from exonerate_b200.engine import span_integrate
    cases = json.load(open(os.path.join(helpers.GOLDEN, "span_cases.json")))
    assert len(cases) >= 30
    for c in cases:
        got = span_integrate(eng, c["src_scores"], c["src_region"], c["dst_region"], c["span"])
        assert got == c["positions"], c["name"]


def test_span_integrate_large_vs_oracle(eng):
    """BSDP-sized and larger regions (48 x 144 and 96 x 288 cells), random and sparse score
    matrices, against the oracle restatement."""
    import test_span_oracle
    from exonerate_b200.engine import span_integrate
    rng = random.Random(5)
    for ql, tl in ((48, 144), (96, 288)):
        n = (ql + 1) * (tl + 1)
        sc = [abi.IMPOSSIBLY_LOW_SCORE if rng.random() < 0.7 else rng.randrange(0, 500) for _ in range(n)]
        case = {"src_scores": sc, "src_region": [100, 5000, ql, tl],
                "dst_region": [100 + ql // 2, 5000 + tl + 300, ql, tl], "span": [0, ql, 30, 200000]}
        assert span_integrate(eng, sc, case["src_region"], case["dst_region"], case["span"]) == \
            test_span_oracle.oracle_span_integrate(case)

