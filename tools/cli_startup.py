#!/usr/bin/env python
"""Where a short `exonerate_b200` run spends its wall time: process + CUDA start-up against the first
device work (tuning aid for profiles/r02_config4.md).  Runs the CLI on inputs that need no DP at all (a random
protein against random DNA: seeding finds nothing), on one planted gene with a cold and a warm cubin cache,
and the reference binary on the same inputs.  usage: python tools/cli_startup.py"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import importlib.util
spec = importlib.util.spec_from_file_location("config4_bench", os.path.join(ROOT, "tools", "config4_bench.py"))
c4 = importlib.util.module_from_spec(spec); spec.loader.exec_module(c4)
base = os.path.join(ROOT, "gpurun_out")
tmp = tempfile.mkdtemp(prefix="startup_", dir=base if os.path.isdir(base) else None)
ours = os.path.join(ROOT, "integration", "_build", "exonerate_b200")
ref = os.path.join(ROOT, "oracle", "_ref", "exonerate_c")
flags = ["--model", "protein2genome", "--exhaustive", "no", "--gappedextension", "no", "--showalignment", "no",
         "--showvulgar", "yes", "--verbose", "0"]


def run(exe, q, t, env=None, label=""):
    t0 = time.perf_counter()
    r = subprocess.run([exe, q, t] + flags, capture_output=True, text=True,
                       env=dict(os.environ, EXONERATE_B200_STATS="1", C4B_TIMING="", **(env or {})))
    dt = time.perf_counter() - t0
    print("%-58s %.2f s  rc=%d  %d alignment(s)" % (label, dt, r.returncode, r.stdout.count("vulgar:")), flush=True)
    return r


for sub, nq, planted in (("none", 1, 0), ("gene", 1, 1)):
    d = os.path.join(tmp, sub); os.makedirs(d)
    q, t = c4.make_inputs(nq, 500, 1000000, planted, 9, d)
    cache = {"C4B_JIT_CACHE_DIR": os.path.join(d, "jit")}
    os.makedirs(cache["C4B_JIT_CACHE_DIR"])
    what = "1 protein x 1 Mbp, %s" % ("no homology (no DP call)" if not planted else "one planted gene")
    run(ref, q, t, label="reference, " + what)
    run(ours, q, t, label="exonerate_b200, " + what + ", no cubin cache")
    run(ours, q, t, env=cache, label="exonerate_b200, " + what + ", cold cubin cache")
    r = run(ours, q, t, env=cache, label="exonerate_b200, " + what + ", warm cubin cache")
    for line in r.stderr.splitlines():
        if line.startswith("exonerate_b200:"):
            print("    " + line[:260])
