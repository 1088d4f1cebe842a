"""Seeded FASTA workloads for the CLI batch tests and the CLI leg of bench.py.

Everything here is generated from a 64-bit LCG written out below, so the files are
identical on every machine and Python version (the goldens under tests/golden/cli_batch/
were produced by running the UNMODIFIED reference on exactly these files,
tests/golden/make_cli_batch_golden.py).  SURVEY.md 8d describes the shapes: the metric
configuration is 1 kbp queries against 100 kbp targets that hold a planted copy of one
query with 15 % edits; config 3 is a 1 kbp cDNA of five exons against 100 kbp genomic."""
import os

MASK = (1 << 64) - 1


class Lcg:
    """Knuth's MMIX LCG; next() = top 32 bits."""

    def __init__(self, seed):
        self.s = (seed * 0x9E3779B97F4A7C15 + 0x1234567) & MASK

    def next(self):
        self.s = (self.s * 6364136223846793005 + 1442695040888963407) & MASK
        return self.s >> 32

    def below(self, n):
        return self.next() % n


def dna(rng, n):
    return "".join("ACGT"[rng.next() >> 30] for _ in range(n))


def mutate(rng, s, per_mille):
    """per-base edits with probability per_mille / 1000: one third each del / ins / sub"""
    out = []
    for ch in s:
        if rng.below(1000) < per_mille:
            kind = rng.below(3)
            if kind == 0:
                continue
            if kind == 1:
                out.append(ch)
                out.append("ACGT"[rng.below(4)])
            else:
                out.append("ACGT"[("ACGT".index(ch) + 1 + rng.below(3)) % 4])
        else:
            out.append(ch)
    return "".join(out)


def write_fasta(path, records):
    with open(path, "w") as f:
        for name, seq in records:
            f.write(">%s\n" % name)
            for k in range(0, len(seq), 70):
                f.write(seq[k:k + 70])
                f.write("\n")


def affine_metric(n_queries, n_targets, qlen=1000, tlen=100000, seed=1):
    """(queries, targets): target j holds a 15 %-edited copy of query j % n_queries"""
    rng = Lcg(seed)
    qs = [("q%d" % k, dna(rng, qlen)) for k in range(n_queries)]
    ts = []
    for j in range(n_targets):
        core = mutate(rng, qs[j % n_queries][1], 150)
        off = rng.below(tlen - len(core))
        ts.append(("t%d" % j, dna(rng, off) + core + dna(rng, tlen - len(core) - off)))
    return qs, ts


def est2genome_metric(n_queries, n_targets, tlen=100000, seed=2):
    """cDNAs of 5 x 200 bp exons; target j = gene of cDNA j % n_queries with GT..AG introns of
    15 kbp, padded to tlen, 2 % edits in the cDNA"""
    rng = Lcg(seed)
    genes = [[dna(rng, 200) for _ in range(5)] for _ in range(n_queries)]
    qs = [("cdna%d" % k, mutate(rng, "".join(ex), 20)) for k, ex in enumerate(genes)]
    ts = []
    for j in range(n_targets):
        ex = genes[j % n_queries]
        body = ex[0]
        for e in ex[1:]:
            body += "GT" + dna(rng, 15000) + "AG" + e
        off = rng.below(tlen - len(body))
        ts.append(("gene%d" % j, dna(rng, off) + body + dna(rng, tlen - len(body) - off)))
    return qs, ts


def write_workload(directory, kind, n_queries, n_targets, *rest, **kw):
    if rest:   # BATCH_COMMANDS entries may carry generator keywords as a fifth element
        kw = dict(rest[0], **kw) if isinstance(rest[0], dict) else kw
    qs, ts = (affine_metric if kind == "affine" else est2genome_metric)(n_queries, n_targets, **kw)
    os.makedirs(directory, exist_ok=True)
    q, t = os.path.join(directory, "q_%s.fa" % kind), os.path.join(directory, "t_%s.fa" % kind)
    write_fasta(q, qs)
    write_fasta(t, ts)
    return q, t


# name -> (kind, n_queries, n_targets, flags): the batch-hook command lines with goldens
COMMON = ["--showalignment", "no", "--showvulgar", "yes", "--showcigar", "yes", "--verbose", "0"]
BATCH_COMMANDS = {
    # the metric shape, the flags SURVEY.md 8d prescribes for parity runs: 24 pairs, one round
    "metric_affine_local": ("affine", 4, 6, ["--model", "affine:local", "--exhaustive", "yes", "--subopt", "no",
                                             "--revcomp", "no", "--score", "0"]),
    # CLI defaults: --subopt yes (sub-optimal series, batched in rounds), both strands, --score 100
    # (20 kbp targets: the reference spends minutes per 100 kbp lattice on a sub-optimal series)
    "metric_affine_local_defaults": ("affine", 3, 4, ["--model", "affine:local", "--exhaustive", "yes"],
                                     {"tlen": 20000}),
    # --bestn changes the threshold as results are submitted: the replay order matters
    "metric_affine_local_bestn": ("affine", 3, 4, ["--model", "affine:local", "--exhaustive", "yes", "--bestn", "1",
                                                   "--subopt", "no"]),
    "metric_est2genome": ("est2genome", 2, 3, ["--model", "est2genome", "--exhaustive", "yes", "--subopt", "no",
                                              "--score", "0"]),
}
