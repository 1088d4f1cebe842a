#!/usr/bin/env python
"""Kernel-only timing of the affine fill for tuning (not a bench number).
usage: python tools/kernel_sweep.py [pairs] [qlen] [tlen] [want_path]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
from bench import make_batch
from exonerate_b200 import Batch, Engine, PairSet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
qlen = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
want_path = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
params = helpers.load_params(); scoring = helpers.load_scoring(params)
model, _ = helpers.load_model("affine_local_dna", params)
queries, targets = make_batch(7, n, qlen, tlen)
pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)])
eng = Engine(0)
b = Batch(eng, model, scoring, pairs, want_path=want_path)
for _ in range(2):
    b.run(); b.last_fill_ms()
ms = []
for _ in range(3):
    b.run(); ms.append(b.last_fill_ms())
res, _ = b.fetch()
chk = sum(res[k].score for k in range(n))
print("lib=%s R=%s pairs=%d %dx%d path=%d fill_ms=%.2f GCUPS=%.1f checksum=%d" % (
    os.path.basename(os.environ.get("C4B_LIB", "default")), os.environ.get("C4B_AFFINE_R", "auto"), n, qlen, tlen,
    want_path, min(ms), pairs.cells / (min(ms) * 1e-3) / 1e9, chk))
