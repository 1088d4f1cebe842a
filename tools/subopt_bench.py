#!/usr/bin/env python
"""SubOpt on the device kernels (tuning aid): iteration 2 of the sub-optimal series -- every lattice
carries the blocked cells of its first path -- against iteration 1, device-resident find_path.
affine:local (int32 BLK kernel) or protein2genome (table-driven path: the systolic specialisation takes
blocked cells as per-strip entries; C4B_JIT_SYSTOLIC=0 shows the thread-per-row kernel it replaced there).
usage: python tools/subopt_bench.py [pairs=2000] [model=affine:local]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from bench import make_batch, make_batch_p2g
from exonerate_b200 import Batch, Engine, PairSet, abi
from exonerate_b200.engine import results_to_list
from exonerate_b200.models import host_model, splice_arrays

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
name = sys.argv[2] if len(sys.argv) > 2 else "affine:local"
params = helpers.load_params(); scoring = helpers.load_scoring(params)
model, _ = host_model(name, query_is_protein=(name == "protein2genome"))
eng = Engine(0)
eng.lib.c4b_engine_set_stream(eng.h, torch.cuda.current_stream().cuda_stream)
splice = None
if name == "protein2genome":
    os.environ.setdefault("C4B_GENERIC_JIT", "1")
    queries, targets = make_batch_p2g(9, n, 450, 20000)
    splice = [splice_arrays(targets[k]) for k in range(n)]
else:
    queries, targets = make_batch(9, n, 1000, 100000)
qs, ts = [queries[k] for k in range(n)], [targets[k] for k in range(n)]


def timed(pairs, label):
    b = Batch(eng, model, scoring, pairs, want_path=True)
    b.run(); b.run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        b.run()
    e1.record(); torch.cuda.synchronize()
    res, ops = b.fetch(ops_capacity=b.ops_needed())
    print("%s: %d pairs, %.0f GCUPS  [%s]" % (label, n, pairs.cells / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e9, b.description),
          flush=True)
    b.close()
    return results_to_list(res, ops, n)


first = timed(PairSet(qs, ts, splice=splice), "iteration 1 (no blocked cells)")
blocked = []
for r in first:   # the match cells of the first path, as SubOpt_add_alignment would block them
    qp, tp, pts = r["region"][0], r["region"][1], []
    for tid, length in r["ops"]:
        tr = model.transitions[tid]
        for _ in range(length):
            if tr.label == abi.LABEL_MATCH:
                pts.append((qp, tp))
            qp += tr.advance_query; tp += tr.advance_target
    blocked.append(pts)
second = timed(PairSet(qs, ts, splice=splice, blocked=blocked), "iteration 2 (first path blocked)")
print("mean score %.1f -> %.1f" % (sum(r["score"] for r in first) / n, sum(r["score"] for r in second) / n))
