"""GPU tier: the BENCH workloads, lattice for lattice, against the UNMODIFIED reference.

oracle/_ref/libc4ref.so is the reference compiled in place (oracle/Makefile) with its own
compiled models; it travels to the GPU box.  For a few lattices of each of bench.py's
workloads at the bench's own shapes -- affine:local and est2genome at 1 kbp x 100 kbp (the
metric shape), protein2genome at 450 aa x 20 kbp -- the GPU's score, alignment region AND
operation list must equal the reference's Optimal_find_path, op for op (these
shapes were only checked against our own kernels or by score)."""
import os
import sys

import pytest

import helpers

sys.path.insert(0, helpers.ROOT)
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from exonerate_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _have_reference():
    from oracle import refdrv
    return refdrv.available()


@pytest.mark.skipif(not _have_reference(), reason="oracle/_ref/libc4ref.so not built (needs the reference sources)")
@pytest.mark.parametrize("model_name,n", [("affine:local", 3), ("est2genome", 3), ("protein2genome", 2)])
def test_bench_shapes_op_for_op_vs_reference(eng, params, scoring, model_name, n, monkeypatch):
    import bench
    from exonerate_b200 import Optimal, PairSet
    from exonerate_b200.models import host_model, splice_arrays
    W = bench.WORKLOADS[model_name]
    qlen, tlen = W.get("qlen", 1000), W.get("tlen", 100000)
    queries, targets = bench.GENERATORS[model_name](4242, n, qlen, tlen)
    model, _ = host_model(model_name, query_is_protein=W.get("query_is_protein", False))
    splice = None
    if model_name != "affine:local":
        splice = [splice_arrays(targets[k]) for k in range(n)]
    if model_name == "protein2genome":
        monkeypatch.setenv("C4B_GENERIC_JIT", "1")   # the kernel the bench measures
    pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)], splice=splice)
    got = Optimal(eng, model, scoring).find_path(pairs)
    pool = bench.CpuPool("reference", n, model_name)   # one reference process per lattice
    try:
        _, want = pool.run(queries, targets, full=True)
    finally:
        pool.close()
    for k in range(n):
        score, region, ops = want[k]
        assert got[k]["score"] == score, (model_name, k)
        assert got[k]["region"] == list(region), (model_name, k)
        assert got[k]["ops"] == [tuple(o) for o in ops], (model_name, k)
        assert score > 500 and len(ops) >= 3
