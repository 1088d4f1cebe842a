#!/usr/bin/env python
"""BASELINE.json configs[4]: GCUPS of affine:local find_path / find_score over a grid of
query x target lengths (kernel-resident timing, CUDA events around Batch.run).  One GPU, or -- launched
with torchrun -- N GPUs sharing every point's FIXED batch (rank r takes pairs r, r+N, ...: strong scaling;
time = max over ranks, GCUPS = all cells / that time).
usage: python tools/sweep_affine.py [total_cells_per_point=2e11] > profiles/...md
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
              tools/sweep_affine.py 4e11"""
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
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
say = print if rank == 0 else (lambda *a, **k: None)
eng = Engine(local)
stream = torch.cuda.current_stream()
eng.lib.c4b_engine_set_stream(eng.h, stream.cuda_stream)
say("| query | target | pairs%s | find_score GCUPS | find_path GCUPS | vs 20 B/cell HBM roofline (%d x 326 GCUPS) |" % (
    " (all %d GPUs)" % world if world > 1 else "", world))
say("|---:|---:|---:|---:|---:|---:|")
only = os.environ.get("SWEEP_ONLY")   # e.g. "2048x1000000,4096x1000000": a subset of the grid (tuning aid)
only = {tuple(int(v) for v in item.split("x")) for item in only.split(",")} if only else None
grid = sorted(only) if only else [(q, t) for q in (128, 512, 1000, 2048, 4096, 16384) for t in (1000, 10000, 100000, 1000000)]
for qlen, tlen in grid:
    if True:
        n = int(max(2, world, min(40000, budget // (qlen * tlen))))
        if n * (qlen + tlen) > 3e9:
            n = int(3e9 // (qlen + tlen))
        queries, targets = make_batch(5, n, qlen, min(tlen, max(tlen, qlen)))
        mine = range(rank, n, world)
        pairs = PairSet([queries[k] for k in mine], [targets[k] for k in mine])
        all_cells = n * queries.shape[1] * targets.shape[1]
        out = []
        for want_path in (False, True):
            try:
                b = Batch(eng, model, scoring, pairs, want_path=want_path)
                b.run(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); b.run(); b.run(); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
                if world > 1:
                    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t[0])
                out.append(all_cells / (ms * 1e-3) / 1e9)
                b.close()
            except Exception as ex:  # e.g. traceback arena beyond the memory budget
                out.append(float("nan"))
                sys.stderr.write("%dx%d path=%d: %s\n" % (qlen, tlen, want_path, ex))
        say("| %d | %d | %d | %.0f | %.0f | %.1fx |" % (qlen, tlen, n, out[0], out[1], out[1] / (326.3 * world)), flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
