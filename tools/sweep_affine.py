#!/usr/bin/env python
"""BASELINE.json configs[4]: GCUPS of affine:local find_path / find_score over a grid of
query x target lengths on one GPU (kernel-resident timing, CUDA events around Batch.run).
usage: python tools/sweep_affine.py [total_cells_per_point=2e11] > profiles/...md"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import helpers
from bench import make_batch
from exonerate_b200 import Batch, Engine, PairSet
from exonerate_b200.models import host_model

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 2e11
params = helpers.load_params(); scoring = helpers.load_scoring(params)
model, _ = host_model("affine:local")
eng = Engine(0)
stream = torch.cuda.current_stream()
eng.lib.c4b_engine_set_stream(eng.h, stream.cuda_stream)
print("| query | target | pairs | find_score GCUPS | find_path GCUPS | vs 20 B/cell HBM roofline (326 GCUPS) |")
print("|---:|---:|---:|---:|---:|---:|")
for qlen in (128, 512, 1000, 2048, 4096, 16384):
    for tlen in (1000, 10000, 100000, 1000000):
        n = int(max(2, min(40000, budget // (qlen * tlen))))
        if n * (qlen + tlen) > 3e9:
            n = int(3e9 // (qlen + tlen))
        queries, targets = make_batch(5, n, qlen, min(tlen, max(tlen, qlen)))
        pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)])
        out = []
        for want_path in (False, True):
            try:
                b = Batch(eng, model, scoring, pairs, want_path=want_path)
                b.run(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); b.run(); b.run(); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
                out.append(pairs.cells / (ms * 1e-3) / 1e9)
                b.close()
            except Exception as ex:  # e.g. traceback arena beyond the memory budget
                out.append(float("nan"))
                sys.stderr.write("%dx%d path=%d: %s\n" % (qlen, tlen, want_path, ex))
        print("| %d | %d | %d | %.0f | %.0f | %.1fx |" % (qlen, tlen, n, out[0], out[1], out[1] / 326.3), flush=True)
eng.close()
