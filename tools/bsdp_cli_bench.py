#!/usr/bin/env python
"""Heuristic (BSDP) protein2genome / est2genome through the CLI: the reference binary with compiled
models on one host core (oracle/_ref/exonerate_c) against the same reference linked with our
viterbi.o / hspset.o (integration/_build/exonerate_b200).  Outputs must be identical; prints wall
times and the binding's own counters (EXONERATE_B200_STATS).  Tuning aid for SURVEY 8a row a13.
usage: python tools/bsdp_cli_bench.py [n_queries] [aa] [target_bp] [model]"""
import os, random, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 4
aa = int(sys.argv[2]) if len(sys.argv) > 2 else 400
tlen = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
model = sys.argv[4] if len(sys.argv) > 4 else "protein2genome"
CODON = {"A": "GCT", "R": "CGT", "N": "AAC", "D": "GAC", "C": "TGC", "Q": "CAG", "E": "GAG", "G": "GGT",
         "H": "CAC", "I": "ATC", "L": "CTG", "K": "AAG", "M": "ATG", "F": "TTC", "P": "CCG", "S": "TCT",
         "T": "ACC", "W": "TGG", "Y": "TAC", "V": "GTT"}
rng = random.Random(2024)
dna = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
queries, genes = [], []
for k in range(nq):
    prot = "".join(rng.choice(list(CODON)) for _ in range(aa))
    cds = "".join(CODON[c] for c in prot)
    cuts = sorted(rng.sample(range(60, len(cds) - 60, 3), 3))
    exons = [cds[a:b] for a, b in zip([0] + cuts, cuts + [len(cds)])]
    gene = ("GT" + dna(rng.randrange(800, 3000)) + "AG").join(exons)
    queries.append(prot if model == "protein2genome" else cds)
    genes.append(gene)
spacer = max(100, (tlen - sum(map(len, genes))) // (nq + 1))
target = dna(spacer) + "".join(g + dna(spacer) for g in genes)
tmp = tempfile.mkdtemp(prefix="bsdp_bench_", dir=os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None)
qf, tf = os.path.join(tmp, "q.fa"), os.path.join(tmp, "t.fa")
with open(qf, "w") as f:
    for k, q in enumerate(queries):
        f.write(">q%d\n%s\n" % (k, q))
with open(tf, "w") as f:
    f.write(">tg\n%s\n" % target)
args = [qf, tf, "--model", model, "--exhaustive", "no", "--gappedextension", "no", "--showalignment", "no",
        "--showvulgar", "yes", "--showcigar", "no", "--verbose", "0"]
outs = {}
for name, exe in (("reference (1 core, compiled models)", os.path.join(ROOT, "oracle", "_ref", "exonerate_c")),
                  ("exonerate_b200", os.path.join(ROOT, "integration", "_build", "exonerate_b200"))):
    if not os.path.exists(exe):
        print(name, "missing:", exe); continue
    t0 = time.perf_counter()
    r = subprocess.run([exe] + args, capture_output=True, text=True, env=dict(os.environ, EXONERATE_B200_STATS="1"))
    dt = time.perf_counter() - t0
    outs[name] = r.stdout
    print("%-40s %.2f s  rc=%d  %d output lines" % (name, dt, r.returncode, len(r.stdout.splitlines())))
    for line in r.stderr.splitlines():
        if line.startswith("exonerate_b200:"):
            print("    " + line)
if len(outs) == 2:
    a, b = outs.values()
    print("outputs identical:", a == b, "| %d queries x %d bp target, model %s" % (nq, len(target), model))
    if a != b:
        print(a[:600]); print(b[:600])
