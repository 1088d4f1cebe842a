#!/usr/bin/env python
"""BASELINE.json configs[1]: 10k synthetic 1 kbp x 1 kbp DNA pairs, affine:local, one GPU
(kernel-resident timing).  usage: python tools/config2_sweep.py [pairs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers
from bench import make_batch
from exonerate_b200 import Batch, Engine, PairSet
from exonerate_b200.models import host_model
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
params = helpers.load_params(); scoring = helpers.load_scoring(params)
model, _ = host_model("affine:local")
eng = Engine(0)
eng.lib.c4b_engine_set_stream(eng.h, torch.cuda.current_stream().cuda_stream)
queries, targets = make_batch(3, n, 1000, 1000)
pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)])
for want_path in (False, True):
    b = Batch(eng, model, scoring, pairs, want_path=want_path)
    b.run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        b.run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res, _ = b.fetch(ops_capacity=n * 2100) if want_path else b.fetch()
    print("TB16=%s %d pairs 1000x1000 path=%d: %.2f ms/step, %.0f GCUPS, checksum %d" % (
        os.environ.get("C4B_AFFINE_TB16", "on"), n, want_path, ms, pairs.cells / (ms * 1e-3) / 1e9,
        sum(res[k].score for k in range(n))), flush=True)
    b.close()
