#!/usr/bin/env python
"""Throughput of c4b_hsp_extend_batch (row a14) next to the CPU oracle port: every 12-mer
match of a 50 kbp query in a 2 Mbp target (+ planted similar regions) as seeds.
usage: python tools/hsp_bench.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
from exonerate_b200 import Engine, HSPset, abi

rng = np.random.default_rng(5)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
q = ACGT[rng.integers(0, 4, 50000)]
t = ACGT[rng.integers(0, 4, 2000000)]
for k in range(200):   # planted, mutated copies of query pieces
    a = int(rng.integers(0, len(q) - 400)); b = int(rng.integers(0, len(t) - 400))
    piece = q[a:a + 400].copy()
    mut = rng.random(400) < 0.08
    piece[mut] = ACGT[rng.integers(0, 4, int(mut.sum()))]
    t[b:b + 400] = piece
qs, ts = bytes(q).decode(), bytes(t).decode()
index = {}
for i in range(len(qs) - 11):
    index.setdefault(qs[i:i + 12], []).append(i)
seeds = [(i, j) for j in range(len(ts) - 11) for i in index.get(ts[j:j + 12], ())]
params = helpers.load_params(); scoring = helpers.load_scoring(params)
param = abi.HspParam(abi.CALC_MATCH_DNA, 12, 30, 75)
eng = Engine(0)
hs = HSPset(eng, scoring, param, qs, ts)
hs.seeds = seeds
hs.extend_all()
# time the C-ABI call alone (host buffers in, host results out), arrays built beforehand
import ctypes as C
n = len(seeds)
sd = (abi.HspSeed * n)(*[abi.HspSeed(a, b) for a, b in seeds])
out = (abi.Hsp * n)()
def call():
    rc = eng.lib.c4b_hsp_extend_batch(eng.h, C.byref(scoring), C.byref(param), hs.q.ctypes.data, len(hs.q), None,
                                      hs.t.ctypes.data, len(hs.t), None, n, sd, out)
    assert rc == 0
call()
t0 = time.perf_counter(); call(); call(); call(); gpu_s = (time.perf_counter() - t0) / 3
ext = out
visited = sum(ext[k].length for k in range(len(seeds)))
n_cpu = min(len(seeds), 20000)
t0 = time.perf_counter(); helpers.oracle_hsp_extend(scoring, param, qs, ts, seeds[:n_cpu]); cpu_s = time.perf_counter() - t0
hsps = hs.finalise()
print("seeds=%d HSPs=%d match-state visits=%d | c4b_hsp_extend_batch end to end (H2D + encode + extend + D2H) %.2f ms = %.2f Mseeds/s | "
      "CPU oracle port (1 core, incl. ctypes) %.2f Mseeds/s" % (len(seeds), len(hsps), visited, gpu_s * 1e3,
                                                               len(seeds) / gpu_s / 1e6, n_cpu / cpu_s / 1e6))
