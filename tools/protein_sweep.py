#!/usr/bin/env python
"""affine:local on proteins (BLOSUM62, int32 kernel with the matrix in shared memory): timing aid.
usage: python tools/protein_sweep.py [pairs] [qlen] [tlen]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
from exonerate_b200 import Batch, Engine, PairSet
from exonerate_b200.models import host_model
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ql = int(sys.argv[2]) if len(sys.argv) > 2 else 400
tl = int(sys.argv[3]) if len(sys.argv) > 3 else 400
params = helpers.load_params(); scoring = helpers.load_scoring(params)
model, _ = host_model("affine:local", query_is_protein=True, target_is_protein=True)
rng = np.random.default_rng(3)
AA = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
queries = AA[rng.integers(0, 20, size=(n, ql))]
targets = AA[rng.integers(0, 20, size=(n, tl))]
targets[:, 50:50 + ql // 2] = queries[:, :ql // 2]
pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)])
eng = Engine(0)
eng.lib.c4b_engine_set_stream(eng.h, torch.cuda.current_stream().cuda_stream)
for want_path in (False, True):
    b = Batch(eng, model, scoring, pairs, want_path=want_path)
    b.run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.run(); b.run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print("protein affine:local kernel=%s pairs=%d %dx%d path=%d %.2f ms GCUPS=%.0f" % (
        b.kernel_name, n, ql, tl, want_path, ms, pairs.cells / (ms * 1e-3) / 1e9), flush=True)
    b.close()
