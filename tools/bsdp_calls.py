#!/usr/bin/env python
"""Which Viterbi_calculate calls does a heuristic (BSDP) run make, and how were they answered?
(EXONERATE_B200_TRACE=1 aggregated by derived-model family; tuning aid for SURVEY 8f row 1.)
usage: python tools/bsdp_calls.py <q.fa> <t.fa> [model]"""
import collections, os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "integration", "_build", "exonerate_b200")
q, t = sys.argv[1], sys.argv[2]
model = sys.argv[3] if len(sys.argv) > 3 else "protein2genome"
env = dict(os.environ, EXONERATE_B200_TRACE="1", EXONERATE_B200_STATS="1")
t0 = time.perf_counter()
r = subprocess.run([exe, q, t, "--model", model, "--exhaustive", "no", "--gappedextension", "no", "--showalignment", "no",
                    "--showvulgar", "yes", "--verbose", "0"], capture_output=True, text=True, env=env)
print("wall %.2f s, %d output lines" % (time.perf_counter() - t0, len(r.stdout.splitlines())))
count = collections.Counter()
cells = collections.Counter()
for line in r.stderr.splitlines():
    m = re.match(r"b200-trace (\w+) mode (\d) \[(.*)\] region (-?\d+) (-?\d+) (-?\d+) (-?\d+) ->", line)
    if not m:
        if line.startswith("exonerate_b200:"):
            print(line)
        continue
    how, mode, name = m.group(1), m.group(2), m.group(3)
    fam = re.split(r"_\d+_|_32_", name)[0][:28]   # mangled names: keep the family prefix
    key = (how, "score" if mode == "0" else "path" if mode == "1" else "region", fam)
    count[key] += 1
    cells[key] += int(m.group(6)) * int(m.group(7))
for key, n in sorted(count.items(), key=lambda kv: -kv[1]):
    print("%-10s %-6s %-30s %8d calls  %12d cells  (%d per call)" % (key + (n, cells[key], cells[key] // max(n, 1))))
