#!/usr/bin/env python
"""Small-batch occupancy of the packed score pass (tuning aid, not a bench number): one
GPU's shard of the fixed 10k-pair batch at N = 8 / 4 / 2 / 1 GPUs, packed two lattices per warp
(C4B_P16_FOLD=0) against folded one lattice per warp (C4B_P16_FOLD=1); C4B_P16_R=16 / 8 turn a
1 kbp query into 2 / 4 pipelined warps per lattice pair (measured slower, profiles/r02_strong_sweep.md).
usage: python tools/strong_sweep.py [model=affine:local]   (run in a fresh process per R: env)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import torch, helpers
    from bench import make_batch
    from exonerate_b200 import Batch, Engine, PairSet
    from exonerate_b200.models import host_model
    n = int(sys.argv[2])
    params = helpers.load_params(); scoring = helpers.load_scoring(params)
    model, _ = host_model("affine:local")
    eng = Engine(0)
    eng.lib.c4b_engine_set_stream(eng.h, torch.cuda.current_stream().cuda_stream)
    queries, targets = make_batch(5, n, 1000, 100000)
    pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)])
    for want_path in (False, True):
        b = Batch(eng, model, scoring, pairs, want_path=want_path)
        b.run(); b.run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            b.run()
        e1.record(); torch.cuda.synchronize()
        print("FOLD=%-4s P16_R=%-4s pairs=%5d path=%d GCUPS=%.0f" % (os.environ.get("C4B_P16_FOLD", "auto"), os.environ.get("C4B_P16_R", "auto"), n, want_path,
              pairs.cells / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e9), flush=True)
        b.close()
    sys.exit(0)

for n in (1250, 2500, 3552, 5000, 7104, 10000):
    for fold in ("0", "1", None):
        env = dict(os.environ)
        if fold: env["C4B_P16_FOLD"] = fold
        else: env.pop("C4B_P16_FOLD", None)
        subprocess.run([sys.executable, __file__, "--one", str(n)], env=env)
