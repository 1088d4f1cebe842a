#!/usr/bin/env python
"""Where the wall time of `exonerate_b200 --exhaustive` goes (tuning aid): the CLI leg of bench.py
with the engine's host timeline (C4B_TIMING) and the binding's counters on stderr.
usage: python tools/cli_timing.py [n_queries] [n_targets]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cli_workload
from bench import make_batch
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 100
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 100
exe = os.path.join(ROOT, "integration", "_build", "exonerate_b200")
flags = ["--model", "affine:local", "--exhaustive", "yes", "--subopt", "no", "--revcomp", "no", "--score", "0"]
with tempfile.TemporaryDirectory() as d:
    qs, ts = make_batch(77, max(nq, nt), 1000, 100000)
    q, t = os.path.join(d, "q.fa"), os.path.join(d, "t.fa")
    cli_workload.write_fasta(q, [("q%d" % k, bytes(qs[k]).decode()) for k in range(nq)])
    cli_workload.write_fasta(t, [("t%d" % k, bytes(ts[k]).decode()) for k in range(nt)])
    for extra in ({}, {"EXONERATE_B200_BATCH_PAIRS": "2000"}):
        env = dict(os.environ, EXONERATE_B200_STATS="1", C4B_TIMING="1", **extra)
        t0 = time.perf_counter()
        got = subprocess.run([exe, q, t] + flags + cli_workload.COMMON, capture_output=True, text=True, env=env)
        wall = time.perf_counter() - t0
        print("== %s: wall %.2f s, %d alignments, %.1f GCUPS" % (extra, wall, got.stdout.count("vulgar:"),
                                                                nq * nt * 1e8 / wall / 1e9))
        print(got.stderr[-6000:])
    # the same FASTA through a run that does (almost) no DP: how long does the reference's own I/O take?
    t0 = time.perf_counter()
    got = subprocess.run([exe, q, t, "--model", "ungapped", "--score", "100000", "--verbose", "0"], capture_output=True,
                         text=True, env=dict(os.environ, EXONERATE_B200_STATS="1"))
    print("== ungapped heuristic run over the same files (FASTA I/O + seeding): %.2f s" % (time.perf_counter() - t0))
    print(got.stderr[-1500:])
