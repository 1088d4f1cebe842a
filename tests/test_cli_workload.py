"""CPU tier: the seeded FASTA workloads behind tests/golden/cli_batch/*.out.

The goldens were produced by the unmodified reference on exactly these sequences
(tests/golden/make_cli_batch_golden.py); the GPU tier regenerates the files, so the generator must
give the same bytes on every machine and Python version: it is a hand-written LCG, pinned here."""
import hashlib
import os

import cli_workload
import helpers

PINNED = {
    "metric_affine_local": "321166492b242e5d",
    "metric_affine_local_bestn": "c861e90e01c4979c",
    "metric_affine_local_defaults": "588088bd3bfbad83",
    "metric_est2genome": "0d61205d86a57d47",
}


def test_workloads_are_pinned():
    assert sorted(PINNED) == sorted(cli_workload.BATCH_COMMANDS)
    for name, entry in cli_workload.BATCH_COMMANDS.items():
        kind, nq, nt = entry[:3]
        kw = entry[4] if len(entry) > 4 else {}
        gen = cli_workload.affine_metric if kind == "affine" else cli_workload.est2genome_metric
        qs, ts = gen(nq, nt, **kw)
        digest = hashlib.sha256("".join(n + s for n, s in qs + ts).encode()).hexdigest()[:16]
        assert digest == PINNED[name], name
        assert os.path.exists(os.path.join(helpers.GOLDEN, "cli_batch", name + ".out")), name


def test_fasta_round_trip(tmp_path):
    q, t = cli_workload.write_workload(str(tmp_path), "affine", 2, 3, tlen=5000)
    recs = open(t).read().split(">")[1:]
    assert len(recs) == 3 and all(len("".join(r.split("\n")[1:])) == 5000 for r in recs)
    assert open(q).read().count(">") == 2


def test_planted_copy_is_found_by_the_oracle(params, scoring):
    """the planted 15 %-edited copy scores far above unrelated flank (what the goldens rely on)"""
    from exonerate_b200 import abi
    model, _ = helpers.load_model("affine_local_dna", params)
    qs, ts = cli_workload.affine_metric(2, 2, qlen=120, tlen=900, seed=5)
    own = helpers.oracle_viterbi(model, scoring, helpers.PairBuf(qs[0][1], ts[0][1]), abi.MODE_FIND_SCORE)["score"]
    other = helpers.oracle_viterbi(model, scoring, helpers.PairBuf(qs[0][1], ts[1][1]), abi.MODE_FIND_SCORE)["score"]
    assert own > 250 > other
