#!/usr/bin/env python
"""Small-batch occupancy of the packed est2genome kernels (tuning aid, not a bench number): one
GPU's shard of the fixed 1k-pair batch (1 kbp cDNA x 100 kbp genomic) at N = 8 / 4 / 2 / 1 GPUs,
rows per lane 16 (512-row sweeps) against 8 (256-row sweeps, four pipelined warps per lattice)
and warps per lattice (C4B_E2G_ROWS / C4B_E2G_WARPS; unset = the library's own choice).
usage: python tools/e2g_small_sweep.py [pairs ...]   (a fresh process per setting: env)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import torch, helpers
    from bench import make_batch_e2g
    from exonerate_b200 import Batch, Engine, PairSet
    from exonerate_b200.models import host_model, splice_arrays
    n = int(sys.argv[2])
    params = helpers.load_params(); scoring = helpers.load_scoring(params)
    model, _ = host_model("est2genome")
    eng = Engine(0)
    eng.lib.c4b_engine_set_stream(eng.h, torch.cuda.current_stream().cuda_stream)
    queries, targets = make_batch_e2g(5, n, 1000, 100000)
    splice = [splice_arrays(targets[k]) for k in range(n)]
    pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)], splice=splice)
    out = []
    for want_path in (False, True):
        b = Batch(eng, model, scoring, pairs, want_path=want_path)
        b.run(); b.run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            b.run()
        e1.record(); torch.cuda.synchronize()
        out.append(pairs.cells / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e9)
        b.close()
    print("ROWS=%-4s WARPS=%-4s pairs=%5d score %.0f path %.0f GCUPS" % (
        os.environ.get("C4B_E2G_ROWS", "auto"), os.environ.get("C4B_E2G_WARPS", "auto"), n, out[0], out[1]), flush=True)
    sys.exit(0)

sizes = [int(a) for a in sys.argv[1:]] or [125, 250, 500, 1000]
for n in sizes:
    for rows, warps in ((None, None), ("16", "1"), ("8", "4"), ("4", "8"), ("4", "4")):
        env = dict(os.environ)
        for key, val in (("C4B_E2G_ROWS", rows), ("C4B_E2G_WARPS", warps)):
            if val: env[key] = val
            else: env.pop(key, None)
        subprocess.run([sys.executable, __file__, "--one", str(n)], env=env)
