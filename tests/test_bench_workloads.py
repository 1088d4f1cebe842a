"""CPU tier: the synthetic workloads bench.py times are deterministic and contain
what they claim -- checked with the oracle on scaled-down instances of the same generators."""
import numpy as np

import bench
import helpers
from exonerate_b200.models import host_model, splice_arrays


def test_generators_are_seeded():
    for gen, args in ((bench.make_batch, (5, 3, 200, 900)), (bench.make_batch_e2g, (5, 3, 200, 3000)),
                      (bench.make_batch_p2g, (5, 3, 60, 2000))):
        a, b = gen(*args), gen(*args)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        c = gen(args[0] + 1, *args[1:])
        assert not np.array_equal(a[1], c[1])


def test_every_workload_has_a_generator_and_defaults():
    assert set(bench.GENERATORS) == set(bench.WORKLOADS)
    for w in bench.WORKLOADS.values():
        assert w["b_alg"] in (20, 80, 104) and w["pairs"] >= 500 and w["cpu_pairs"] >= 1


def test_protein2genome_workload_plants_a_spliced_gene(params, scoring):
    """make_batch_p2g at about 1/3 scale: the oracle's protein2genome path must cross introns
    (5'ss / intron / 3'ss labels, one set per intron; exons of a few residues may be skipped) and
    score well above an unrelated pair."""
    model, _ = host_model("protein2genome", query_is_protein=True)
    queries, targets = bench.make_batch_p2g(11, 2, 160, 4000)
    q, t = bytes(queries[0]).decode(), bytes(targets[0]).decode()
    got = helpers.oracle_find_path(model, scoring, helpers.PairBuf(q, t, splice=splice_arrays(t)),
                                   region_threshold_cells=0, max_ops=len(q) + len(t) + 8)
    other = helpers.oracle_find_path(model, scoring,
                                     helpers.PairBuf(q, bytes(targets[1]).decode(),
                                                     splice=splice_arrays(bytes(targets[1]).decode())),
                                     region_threshold_cells=0, max_ops=len(q) + len(t) + 8)
    vulgar = helpers.format_ops(model, got["ops"], "vulgar")
    introns = vulgar.count(" I ")
    assert 2 <= introns <= 3 and vulgar.count(" 5 ") == introns and vulgar.count(" 3 ") == introns, vulgar
    assert got["score"] > 3 * max(other["score"], 1)
    assert got["region"][1] - got["region"][0] >= 120  # most of the protein aligned
