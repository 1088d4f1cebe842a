#!/usr/bin/env python
"""BASELINE.json configs[3]: protein2genome, 500 aa queries x 1 Mbp genomic, heuristic mode (BSDP bounded
regions), sharded by COMPARISON over GPUs and host processes ("BSDP: shard by comparison, not by
region").  The reference CLI is single-threaded and its own --querychunkid / --querychunktotal split a
query file between processes; each process here is `exonerate_b200` (the reference objects + our
viterbi.o / hspset / hpair / gam bindings) pinned to one GPU with EXONERATE_B200_DEVICE.  The same chunking
of the unmodified reference binary (compiled models) on the same host cores is the CPU baseline, and the
concatenated outputs must be byte-identical.

Reports wall time end to end and, from the binding's counters (EXONERATE_B200_STATS), the DP-only part: the
host-visible time of every Viterbi_calculate / batched region fill.  What is left is the reference's own
host code (FASTA, seeding, SAR graph, printing), which north_star keeps as-is.

Several processes on ONE GPU time-slice it (every batch's synchronise waits for the other contexts' slices);
--mps starts the CUDA MPS daemon for the run so their kernels overlap instead.  The model-specialised kernels
are compiled once into C4B_JIT_CACHE_DIR by an untimed one-query run (the counterpart of the reference's
bootstrapper step, reported separately).

usage: python tools/config4_bench.py [--queries 1000] [--aa 500] [--target 1000000] [--planted 100]
                                     [--gpus 1] [--procs 8] [--mps] [--no-reference]"""
import argparse, os, re, shutil, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODON = {"A": "GCT", "R": "CGT", "N": "AAC", "D": "GAC", "C": "TGC", "Q": "CAG", "E": "GAG", "G": "GGT",
         "H": "CAC", "I": "ATC", "L": "CTG", "K": "AAG", "M": "ATG", "F": "TTC", "P": "CCG", "S": "TCT",
         "T": "ACC", "W": "TGG", "Y": "TAC", "V": "GTT"}


def make_inputs(nq, aa, tlen, planted, seed, out_dir):
    """nq random proteins; the first `planted` of them have a 4-exon gene (back-translated, 2 % substitutions,
    GT..AG introns of 0.8-3 kbp) in the target, alternating strands; the rest meet only random sequence."""
    rng = np.random.default_rng(seed)
    letters = np.array(list(CODON))
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    prots = ["".join(letters[rng.integers(0, 20, aa)]) for _ in range(nq)]
    genes = []
    for k in range(planted):
        cds = np.frombuffer("".join(CODON[c] for c in prots[k]).encode(), dtype=np.uint8).copy()
        sub = rng.random(cds.size) < 0.02
        cds[sub] = acgt[rng.integers(0, 4, int(sub.sum()))]
        cds = cds.tobytes()
        cuts = sorted(3 * int(c) for c in rng.choice(np.arange(20, aa - 20), 3, replace=False))
        exons = [cds[a:b] for a, b in zip([0] + cuts, cuts + [len(cds)])]
        gene = b""
        for e, ex in enumerate(exons):
            gene += ex
            if e + 1 < len(exons):
                gene += b"GT" + acgt[rng.integers(0, 4, int(rng.integers(800, 3000)))].tobytes() + b"AG"
        if k & 1:
            gene = gene.translate(comp)[::-1]
        genes.append(gene)
    spare = tlen - sum(map(len, genes))
    if spare < 0:
        raise SystemExit("%d planted genes do not fit %d bp" % (planted, tlen))
    gaps = rng.multinomial(spare, np.ones(planted + 1) / (planted + 1))
    parts = []
    for k in range(planted + 1):
        parts.append(acgt[rng.integers(0, 4, int(gaps[k]))].tobytes())
        if k < planted:
            parts.append(genes[k])
    target = b"".join(parts)
    order = rng.permutation(nq)   # planted queries spread over the chunks
    qf, tf = os.path.join(out_dir, "q.fa"), os.path.join(out_dir, "t.fa")
    with open(qf, "w") as f:
        for k in order:
            f.write(">q%d\n%s\n" % (k, prots[k]))
    with open(tf, "wb") as f:
        f.write(b">tg\n" + target + b"\n")
    return qf, tf


def run_chunks(exe, args, procs, gpus):
    """`procs` processes, chunk c of procs each, process c on GPU c % gpus; returns wall, outputs, stderrs"""
    t0 = time.perf_counter()
    ps = []
    for c in range(procs):
        env = dict(os.environ, EXONERATE_B200_STATS="1")
        if gpus:   # one visible device per process: the driver initialises only that one
            env["CUDA_VISIBLE_DEVICES"] = str(c % gpus)
            env["EXONERATE_B200_DEVICE"] = "0"
        ps.append(subprocess.Popen([exe] + args + ["--querychunkid", str(c + 1), "--querychunktotal", str(procs)],
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate() for p in ps]
    wall = time.perf_counter() - t0
    if any(p.returncode for p in ps):
        raise SystemExit("%s failed: %s" % (exe, [o[1][-400:] for p, o in zip(ps, outs) if p.returncode]))
    return wall, [o[0] for o in outs], [o[1] for o in outs]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=1000)
    ap.add_argument("--aa", type=int, default=500)
    ap.add_argument("--target", type=int, default=1000000)
    ap.add_argument("--planted", type=int, default=100)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--mps", action="store_true")
    ap.add_argument("--no-jit-cache", action="store_true")
    a = ap.parse_args()
    base = os.path.join(ROOT, "gpurun_out")
    tmp = tempfile.mkdtemp(prefix="config4_", dir=base if os.path.isdir(base) else None)
    qf, tf = make_inputs(a.queries, a.aa, a.target, a.planted, a.seed, tmp)
    args = [qf, tf, "--model", "protein2genome", "--exhaustive", "no", "--gappedextension", "no",
            "--showalignment", "no", "--showvulgar", "yes", "--showcigar", "no", "--verbose", "0"]
    print("config 4: %d proteins of %d aa x one %d bp genomic target (both strands = %d comparisons), %d with a planted "
          "4-exon gene; %d process(es) on %d GPU(s), host cores %d" % (
              a.queries, a.aa, a.target, 2 * a.queries, a.planted, a.procs, a.gpus, os.cpu_count() or 0))
    ours = os.path.join(ROOT, "integration", "_build", "exonerate_b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "exonerate_c")
    if not a.no_jit_cache:
        os.environ["C4B_JIT_CACHE_DIR"] = os.path.join(tmp, "jit")
        os.makedirs(os.environ["C4B_JIT_CACHE_DIR"], exist_ok=True)
        wq = os.path.join(tmp, "warm.fa")
        with open(qf) as f, open(wq, "w") as g:   # queries with a planted gene reach every derived model
            lines = f.read().split(">")[1:]
            g.write("".join(">" + x for x in lines if int(x.split()[0][1:]) < max(1, min(4, a.planted))))
        t0 = time.perf_counter()
        subprocess.run([ours, wq] + args[1:], capture_output=True, env=dict(os.environ, EXONERATE_B200_DEVICE="0"))
        print("kernel specialisation (one untimed run, %d cubins kept in C4B_JIT_CACHE_DIR): %.1f s" % (
            len(os.listdir(os.environ["C4B_JIT_CACHE_DIR"])), time.perf_counter() - t0))
    mps = None
    if a.mps or a.procs > a.gpus:   # several contexts on one GPU: let their kernels overlap
        ctl = shutil.which("nvidia-cuda-mps-control")
        if ctl:
            os.environ["CUDA_MPS_PIPE_DIRECTORY"] = os.path.join(tmp, "mps")
            os.environ["CUDA_MPS_LOG_DIRECTORY"] = os.path.join(tmp, "mps_log")
            os.makedirs(os.environ["CUDA_MPS_PIPE_DIRECTORY"], exist_ok=True)
            os.makedirs(os.environ["CUDA_MPS_LOG_DIRECTORY"], exist_ok=True)
            mps = subprocess.run([ctl, "-d"], capture_output=True, text=True)
            print("CUDA MPS daemon: rc %d %s" % (mps.returncode, (mps.stderr or mps.stdout).strip()[:200]))
            time.sleep(1.0)
        else:
            print("CUDA MPS daemon: nvidia-cuda-mps-control not found, processes time-slice the GPU")
    try:
        wall, outs, errs = run_chunks(ours, args, a.procs, a.gpus)
    finally:
        if mps is not None and mps.returncode == 0:
            subprocess.run([shutil.which("nvidia-cuda-mps-control")], input="quit\n", capture_output=True, text=True)
            for k in ("CUDA_MPS_PIPE_DIRECTORY", "CUDA_MPS_LOG_DIRECTORY"):
                os.environ.pop(k, None)
    n_aln = sum(o.count("vulgar:") for o in outs)
    dp = fills = batches = 0.0
    for e in errs:
        for m in re.finditer(r"Viterbi_calculate calls (\d+) \(([\d.]+) s", e):
            dp += float(m.group(2))
        for m in re.finditer(r"fills prefetched (\d+) in (\d+) batch", e):
            fills += int(m.group(1)); batches += int(m.group(2))
    print("exonerate_b200: wall %.2f s end to end (%.1f comparisons/s), %d alignments; DP-only (host-visible time of all "
          "Viterbi_calculate calls, summed over processes) %.2f s; %d region fills in %d device batches" % (
              wall, 2 * a.queries / wall, n_aln, dp, fills, batches))
    for e in errs[:1]:
        for line in e.splitlines():
            if line.startswith("exonerate_b200:"):
                print("    [process 1] " + line)
    if not a.no_reference and os.path.exists(ref):
        rwall, routs, _ = run_chunks(ref, args, a.procs, 0)
        print("reference (compiled models), same %d-way split on the host cores: wall %.2f s (%.1f comparisons/s)" % (
            a.procs, rwall, 2 * a.queries / rwall))
        print("outputs identical: %s | speed-up end to end %.2fx" % (outs == routs, rwall / wall))
        if outs != routs:
            for c, (x, y) in enumerate(zip(outs, routs)):
                if x != y:
                    print("first differing chunk %d:\n%s\n---\n%s" % (c + 1, x[:500], y[:500]))
                    break
            return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
