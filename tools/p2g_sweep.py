#!/usr/bin/env python
"""protein2genome / coding2coding on the table-driven kernel: a first timing for DESIGN.md
(these models have no specialised kernel yet).  usage: python tools/p2g_sweep.py [pairs] [aa] [tlen]"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers
from exonerate_b200 import Batch, Engine, PairSet
from exonerate_b200.models import host_model, splice_arrays

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
aa = int(sys.argv[2]) if len(sys.argv) > 2 else 500
tlen = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
params = helpers.load_params(); scoring = helpers.load_scoring(params)
eng = Engine(0)
eng.lib.c4b_engine_set_stream(eng.h, torch.cuda.current_stream().cuda_stream)
rng = random.Random(1)
for name, qprot in (("protein2genome", True), ("coding2coding", False)):
    model, _ = host_model(name, query_is_protein=qprot)
    qs, ts, sp = [], [], []
    for k in range(n):
        if qprot:
            q = helpers.rand_dna(rng, aa, helpers.PROTEIN_ALPHABET)
        else:
            q = helpers.rand_dna(rng, 3 * aa)
        t = helpers.rand_dna(rng, tlen)
        qs.append(q); ts.append(t); sp.append(splice_arrays(t) if name == "protein2genome" else None)
    pairs = PairSet(qs, ts, splice=sp)
    for want_path in (False, True):
        b = Batch(eng, model, scoring, pairs, want_path=want_path)
        b.run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.run(); e1.record(); torch.cuda.synchronize()
        print("%s kernel=%s pairs=%d %dx%d path=%d GCUPS=%.2f" % (name, b.kernel_name, n, len(qs[0]), tlen, want_path,
              pairs.cells / (e0.elapsed_time(e1) * 1e-3) / 1e9), flush=True)
        b.close()
