#!/usr/bin/env python
"""Golden CLI outputs: the UNMODIFIED reference binary (oracle/_ref/exonerate_c,
compiled models) on small seeded FASTA files, `--exhaustive yes`.  The GPU tier
runs the same command lines through integration/_build/exonerate_b200 (the
reference with OUR viterbi.o linked in) and requires byte-identical stdout.

Runs only in the build container.  Usage: python tests/golden/make_cli_golden.py
"""
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

CLI = os.path.join(HERE, "cli")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "exonerate_c")

COMMON = ["--showalignment", "no", "--showvulgar", "yes", "--showcigar", "yes", "--verbose", "0"]
# name -> (query fasta, target fasta, extra flags)
COMMANDS = {
    "affine_local_defaults": ("q_dna.fa", "t_dna.fa", ["--model", "affine:local", "--exhaustive", "yes"]),
    "affine_local_nosubopt": ("q_dna.fa", "t_dna.fa", ["--model", "affine:local", "--exhaustive", "yes",
                                                      "--subopt", "no", "--score", "0"]),
    "affine_global": ("q_dna.fa", "t_dna.fa", ["--model", "affine:global", "--exhaustive", "yes",
                                               "--subopt", "no", "--revcomp", "no", "--score", "-100000"]),
    "affine_bestfit": ("q_dna.fa", "t_dna.fa", ["--model", "affine:bestfit", "--exhaustive", "yes",
                                                "--subopt", "no", "--revcomp", "no", "--score", "-100000"]),
    "affine_local_gaps": ("q_dna.fa", "t_dna.fa", ["--model", "affine:local", "--exhaustive", "yes",
                                                   "--gapopen", "-7", "--gapextend", "-2", "--subopt", "no"]),
    "affine_local_protein": ("q_prot.fa", "t_prot.fa", ["--model", "affine:local", "--exhaustive", "yes"]),
    "est2genome": ("q_cdna.fa", "t_gene.fa", ["--model", "est2genome", "--exhaustive", "yes"]),
    "protein2genome": ("q_prot.fa", "t_gene_p.fa", ["--model", "protein2genome", "--exhaustive", "yes",
                                                    "--subopt", "no"]),
    "coding2coding": ("q_cds.fa", "t_cds.fa", ["--model", "coding2coding", "--exhaustive", "yes", "--subopt", "no"]),
    # heuristic (BSDP) mode: HSP seeding + SAR terminal / join fills on derived models and the
    # bound fills of Heuristic_create -- every Viterbi call still goes through our viterbi.o
    "bsdp_affine_local_dna": ("q_dna.fa", "t_dna.fa", ["--model", "affine:local", "--exhaustive", "no",
                                                       "--gappedextension", "no"]),
    "bsdp_affine_local_protein": ("q_prot.fa", "t_prot.fa", ["--model", "affine:local", "--exhaustive", "no",
                                                             "--gappedextension", "no"]),
    # ... and the span models of the spliced / codon models (cell_start / cell_end callbacks
    # evaluated by the binding around the device fill, c4b_viterbi_calculate_cells)
    "bsdp_est2genome": ("q_cdna.fa", "t_gene.fa", ["--model", "est2genome", "--exhaustive", "no",
                                                   "--gappedextension", "no"]),
    "bsdp_protein2genome": ("q_prot.fa", "t_gene_p.fa", ["--model", "protein2genome", "--exhaustive", "no",
                                                         "--gappedextension", "no"]),
    "bsdp_coding2coding": ("q_cds.fa", "t_cds.fa", ["--model", "coding2coding", "--exhaustive", "no",
                                                    "--gappedextension", "no"]),
    # default heuristic mode (--gappedextension yes): HSPs seeded on the device through the
    # hspset binding, gapped extension by the reference's own SDP code (out of scope, CPU)
    "sdp_affine_local_dna": ("q_dna.fa", "t_dna.fa", ["--model", "affine:local"]),
    "sdp_est2genome": ("q_cdna.fa", "t_gene.fa", ["--model", "est2genome"]),
    "ungapped_dna": ("q_dna.fa", "t_dna.fa", ["--model", "ungapped"]),
    "ryo": ("q_dna.fa", "t_dna.fa", ["--model", "affine:local", "--exhaustive", "yes",
                                                   "--subopt", "no", "--ryo", "%qi %ti %s %pi %em\\n"]),
}

CODON = {"A": "GCT", "R": "CGT", "N": "AAT", "D": "GAT", "C": "TGT", "Q": "CAA", "E": "GAA", "G": "GGT",
         "H": "CAT", "I": "ATT", "L": "CTG", "K": "AAA", "M": "ATG", "F": "TTT", "P": "CCT", "S": "TCT",
         "T": "ACT", "W": "TGG", "Y": "TAT", "V": "GTT"}
COMP = str.maketrans("ACGT", "TGCA")


def fasta(path, records):
    with open(path, "w") as f:
        for name, seq in records:
            f.write(">%s\n" % name)
            for k in range(0, len(seq), 60):
                f.write(seq[k:k + 60] + "\n")


def main():
    os.makedirs(CLI, exist_ok=True)
    rng = random.Random(20260101)
    # DNA: three queries; targets hold mutated copies, one on the reverse strand, one twice
    qs = [helpers.rand_dna(rng, n) for n in (180, 260, 333)]
    t0 = helpers.rand_dna(rng, 300) + helpers.mutate(rng, qs[0], 0.1) + helpers.rand_dna(rng, 500)
    t1 = helpers.rand_dna(rng, 150) + helpers.mutate(rng, qs[1], 0.15).translate(COMP)[::-1] + helpers.rand_dna(rng, 200)
    t2 = (helpers.rand_dna(rng, 100) + helpers.mutate(rng, qs[2], 0.05) + helpers.rand_dna(rng, 400) +
          helpers.mutate(rng, qs[2], 0.2) + helpers.rand_dna(rng, 100))
    fasta(os.path.join(CLI, "q_dna.fa"), [("q%d" % k, s) for k, s in enumerate(qs)])
    fasta(os.path.join(CLI, "t_dna.fa"), [("t0", t0), ("t1", t1), ("t2", t2)])
    # proteins
    ps = [helpers.rand_dna(rng, n, helpers.PROTEIN_ALPHABET) for n in (60, 95)]
    tp = [helpers.rand_dna(rng, 40, helpers.PROTEIN_ALPHABET) + helpers.mutate(rng, p, 0.15, helpers.PROTEIN_ALPHABET) +
          helpers.rand_dna(rng, 30, helpers.PROTEIN_ALPHABET) for p in ps]
    fasta(os.path.join(CLI, "q_prot.fa"), [("p%d" % k, s) for k, s in enumerate(ps)])
    fasta(os.path.join(CLI, "t_prot.fa"), [("tp%d" % k, s) for k, s in enumerate(tp)])
    # a gene: cDNA of 3 exons, genomic with GT..AG introns (and its reverse complement)
    exons = [helpers.rand_dna(rng, n) for n in (90, 70, 110)]
    cdna = "".join(exons)
    gene = (helpers.rand_dna(rng, 200) + exons[0] + "GT" + helpers.rand_dna(rng, 150) + "AG" + exons[1] +
            "GT" + helpers.rand_dna(rng, 240) + "AG" + exons[2] + helpers.rand_dna(rng, 180))
    fasta(os.path.join(CLI, "q_cdna.fa"), [("cdna", helpers.mutate(rng, cdna, 0.02))])
    fasta(os.path.join(CLI, "t_gene.fa"), [("gene", gene), ("gene_rc", gene.translate(COMP)[::-1])])
    # a protein-coding gene for protein2genome: exons are back-translated protein pieces
    cds = "".join(CODON[a] for a in ps[0])
    cut1, cut2 = 60, 127  # phase 0 and phase 1 introns
    gene_p = (helpers.rand_dna(rng, 120) + cds[:cut1] + "GT" + helpers.rand_dna(rng, 90) + "AG" + cds[cut1:cut2] +
              "GT" + helpers.rand_dna(rng, 130) + "AG" + cds[cut2:] + helpers.rand_dna(rng, 100))
    fasta(os.path.join(CLI, "t_gene_p.fa"), [("gene_p", gene_p)])
    # coding vs coding
    cds2 = helpers.mutate(rng, cds, 0.04)
    fasta(os.path.join(CLI, "q_cds.fa"), [("cds", cds)])
    fasta(os.path.join(CLI, "t_cds.fa"), [("cds2", helpers.rand_dna(rng, 30) + cds2 + helpers.rand_dna(rng, 33))])

    manifest = {}
    for name, (q, t, flags) in COMMANDS.items():
        args = [q, t] + flags + COMMON
        out = subprocess.run([REF_BIN] + args, cwd=CLI, capture_output=True, text=True, check=True).stdout
        with open(os.path.join(CLI, name + ".out"), "w") as f:
            f.write(out)
        manifest[name] = args
        print(name, len(out.splitlines()), "lines")
    with open(os.path.join(CLI, "commands.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
