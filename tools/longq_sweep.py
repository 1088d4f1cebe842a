"""Long-query corner of BASELINE.json configs[4] (tuning aid): GCUPS with 1 warp per lattice
(C4B_AFFINE_WARPS=1) against the pipelined sweeps (default)."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, helpers
from bench import make_batch
from exonerate_b200 import Batch, Engine, PairSet
from exonerate_b200.models import host_model
params = helpers.load_params(); scoring = helpers.load_scoring(params)
model, _ = host_model("affine:local")
eng = Engine(0)
eng.lib.c4b_engine_set_stream(eng.h, torch.cuda.current_stream().cuda_stream)   # events below see the engine's work
for qlen, tlen, n in ((16384, 100000, 122), (4096, 100000, 488), (2048, 100000, 976), (16384, 1000000, 12), (16384, 10000, 1220)):
    queries, targets = make_batch(5, n, qlen, tlen)
    pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)])
    for want_path in (False, True):
        b = Batch(eng, model, scoring, pairs, want_path=want_path)
        b.run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.run(); e1.record(); torch.cuda.synchronize()
        print("W=%s %dx%d n=%d path=%d GCUPS=%.0f" % (os.environ.get("C4B_AFFINE_WARPS", "auto"), qlen, tlen, n, want_path, pairs.cells / (e0.elapsed_time(e1) * 1e-3) / 1e9), flush=True)
        b.close()
