#!/usr/bin/env python
"""The windowed PATH route of the table-driven kernel (column checkpoints + window refills,
DESIGN.md 4.3d JIT_SYS_WIN) at bench scale: protein2genome lattices of bench.py's generator
(planted 4-exon genes), find_path once with the whole record (the default route) and once forced
through windows (C4B_GENERIC_TB_BUDGET_KB=1), results compared op for op, both timed.
usage: python tools/window_check.py [pairs=128] [aa=450] [tlen=100000] [window_cols=4096]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers
from bench import make_batch_p2g
from exonerate_b200 import Engine, Optimal, PairSet
from exonerate_b200.models import host_model, splice_arrays

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
aa = int(sys.argv[2]) if len(sys.argv) > 2 else 450
tlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
wcols = sys.argv[4] if len(sys.argv) > 4 else "4096"
os.environ["C4B_GENERIC_JIT"] = "1"
params = helpers.load_params(); scoring = helpers.load_scoring(params)
model, _ = host_model("protein2genome", query_is_protein=True)
eng = Engine(0)
queries, targets = make_batch_p2g(11, n, aa, tlen)
pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)],
                splice=[splice_arrays(targets[k]) for k in range(n)])
opt = Optimal(eng, model, scoring)
out = {}
for route, env in (("whole record", {}), ("column windows", {"C4B_GENERIC_TB_BUDGET_KB": "1", "C4B_GENERIC_WINDOW_COLS": wcols})):
    for k in ("C4B_GENERIC_TB_BUDGET_KB", "C4B_GENERIC_WINDOW_COLS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    opt.find_path(pairs)   # (first call compiles the kernels of the route)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out[route] = opt.find_path(pairs)
    dt = time.perf_counter() - t0
    print("%-15s %d lattices of %d aa x %d bp: %.3f s = %.1f GCUPS (find_path through the C ABI, host buffers)" % (
        route, n, aa, tlen, dt, pairs.cells / dt / 1e9), flush=True)
same = out["whole record"] == out["column windows"]
spans = [r["region"] for r in out["whole record"][:3]]
print("results identical (score, region, ops of every lattice):", same, "| first regions:", spans)
sys.exit(0 if same else 1)
