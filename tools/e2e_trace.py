#!/usr/bin/env python
"""Host-side timeline of c4b_find_path_batch on the metric workload (C4B_TIMING=1)."""
import os, sys, time
os.environ["C4B_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
from bench import make_batch
from exonerate_b200 import Engine, Optimal, PairSet, abi
from exonerate_b200.models import host_model
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
params = helpers.load_params(); scoring = helpers.load_scoring(params)
model, _ = host_model("affine:local")
queries, targets = make_batch(1000, n, 1000, 100000)
pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)])
eng = Engine(0); opt = Optimal(eng, model, scoring)
out = ((abi.Result * n)(), np.empty(2 * n * 4096, dtype=np.int32))
for it in range(3):
    t0 = time.perf_counter(); opt.find_path_raw(pairs, out=out); dt = time.perf_counter() - t0
    sys.stderr.write("[python] call %d: %.2f ms\n" % (it, dt * 1e3))
