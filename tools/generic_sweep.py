#!/usr/bin/env python
"""est2genome on the table-driven kernel: timing + a parity spot check (tuning aid).
usage: python tools/generic_sweep.py [pairs] [qlen] [tlen] [want_path]"""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
from exonerate_b200 import Batch, Engine, PairSet, abi
from exonerate_b200.models import host_model, splice_arrays

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
qlen = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
want_path = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True


def gene(seed, qlen, tlen, n_exons=5):
    rng = random.Random(seed)
    exon = qlen // n_exons
    exons = [helpers.rand_dna(rng, exon) for _ in range(n_exons)]
    q = "".join(exons)
    intron = max(40, (tlen * 3 // 4 - qlen) // max(1, n_exons - 1))
    body = ""
    for k, e in enumerate(exons):
        body += helpers.mutate(rng, e, 0.02)
        if k + 1 < n_exons:
            body += "GT" + helpers.rand_dna(rng, intron) + "AG"
    pad = max(0, tlen - len(body))
    left = rng.randrange(0, pad + 1)
    t = helpers.rand_dna(rng, left) + body + helpers.rand_dna(rng, pad - left)
    return q, t[:max(tlen, len(body))]


params = helpers.load_params(); scoring = helpers.load_scoring(params)
model, _ = host_model("est2genome")
qs, ts, sp = [], [], []
for k in range(n):
    q, t = gene(100 + k, qlen, tlen)
    qs.append(q); ts.append(t); sp.append(splice_arrays(t))
pairs = PairSet(qs, ts, splice=sp)
eng = Engine(0)
b = Batch(eng, model, scoring, pairs, want_path=want_path)
t0 = time.perf_counter(); b.run(); ms0 = b.last_fill_ms()
b.run(); ms = b.last_fill_ms()
res, ops = b.fetch(ops_capacity=n * 4096)
print("kernel=%s pairs=%d %dx%d path=%d fill_ms=%.1f GCUPS=%.3f score0=%d n_ops0=%d" % (
    b.kernel_name, n, qlen, tlen, want_path, ms, pairs.cells / (ms * 1e-3) / 1e9, res[0].score, res[0].n_ops))
if qlen * tlen <= 3e7:
    want = helpers.oracle_find_path(model, scoring, helpers.PairBuf(qs[0], ts[0], splice=sp[0]),
                                    region_threshold_cells=0, max_ops=qlen + tlen)
    o = int(res[0].ops_offset)
    got_ops = [(int(ops[2 * (o + i)]), int(ops[2 * (o + i) + 1])) for i in range(res[0].n_ops)]
    print("oracle parity:", res[0].score == want["score"] and got_ops == want["ops"], want["score"])
    print(helpers.format_ops(model, got_ops, "vulgar")[:200])
