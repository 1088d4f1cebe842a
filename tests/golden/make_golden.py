#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED
reference (oracle/_ref/libc4ref.so, built by `make -C oracle ref`).

Runs only in the build container (needs /root/reference at build time); the
fixtures it writes are committed so that the CPU and GPU test tiers never need
the reference.  Usage:  python tests/golden/make_golden.py

Writes:
  models/<name>.txt        closed C4_Model dumps (transition order = contract)
  scoring.json             Submat / Translate tables and penalties
  cases_<name>.json        inputs + reference score / region / ops / vulgar / cigar
  splice_<name>.npz        per-target splice-site int arrays for intron models
"""
import ctypes as C
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refdrv  # noqa: E402
import helpers  # noqa: E402

# (fixture name, reference model name, query protein?, target protein?)
MODELS = [
    ("affine_local_dna", "affine:local", 0, 0),
    ("affine_global_dna", "affine:global", 0, 0),
    ("affine_bestfit_dna", "affine:bestfit", 0, 0),
    ("affine_overlap_dna", "affine:overlap", 0, 0),
    ("affine_local_protein", "affine:local", 1, 1),
    ("affine_global_protein", "affine:global", 1, 1),
    ("affine_bestfit_protein", "affine:bestfit", 1, 1),
    ("affine_overlap_protein", "affine:overlap", 1, 1),
    ("ungapped_dna", "ungapped", 0, 0),
    ("est2genome", "est2genome", 0, 0),
    ("protein2genome", "protein2genome", 1, 0),
    ("coding2coding", "coding2coding", 0, 0),
]

# the reference's own KAT inputs (src/model/affine.test.c:34-40,
# est2genome.test.c:25-36, protein2genome.test.c, coding2coding.test.c)
KAT_AFFINE_Q = "MEEPQSDPSVEPPLSQETFSDLWKLL"
KAT_AFFINE_T = "PENNVLSPLPSQAMDDLMLSPDDIEQWFTEDPGPEHSCETFDIWKWCPIECDFLNVISEPNEPIPSQ"
KAT_E2G_Q = "CGATCGATCGNATCGATCGATC" "CATCTATCTAGCGAGCGATCTA"
KAT_E2G_T = ("CGATCGATCGATCGATCGATC" "GTNNNNNNNNNNNNNNNNNNNN" + "N" * 47 + "N" * 47 + "N" * 47 +
             "NNNNNNNNNNNNNNNNNNNNNNNNNNNAG" "CATCTATCTANNNGCGAGCGATCTA")


def read_kat_strings(path):
    """Pull the (query, target) string literals out of the Sequence_create()
    calls of a reference test source (inputs only; read at generation time)."""
    import re
    src = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    seqs = []
    for m in re.finditer(r'Sequence_create\(\s*"\w+"\s*,\s*NULL\s*,\s*((?:"[^"]*"\s*)+),', src):
        seqs.append("".join(re.findall(r'"([^"]*)"', m.group(1))))
    assert len(seqs) >= 2, path
    return {"query_seq": seqs[0], "target_seq": seqs[1]}


def est_genome_pair(seed, n_exons=3, exon_len=60, intron_len=90, flank=40):
    rng = random.Random(seed)
    exons = [helpers.rand_dna(rng, exon_len) for _ in range(n_exons)]
    q = "".join(exons)
    t = helpers.rand_dna(rng, flank)
    for k, e in enumerate(exons):
        t += helpers.mutate(rng, e, 0.03)
        if k + 1 < n_exons:
            t += "GT" + helpers.rand_dna(rng, intron_len) + "AG"
    t += helpers.rand_dna(rng, flank)
    return q, t


def main():
    os.makedirs(os.path.join(HERE, "models"), exist_ok=True)

    def run(lib):
        # ---- scoring tables -------------------------------------------------
        params = {}
        for key, protein in (("dna", 0), ("protein", 1)):
            mat = (C.c_int * 576)()
            idx = (C.c_ubyte * 256)()
            lib.c4ref_get_submat(protein, mat, idx)
            params[key + "_matrix"] = list(mat)
            params[key + "_index"] = bytes(idx).hex()
        nt2d = (C.c_ubyte * 256)()
        caa = (C.c_ubyte * 4096)()
        lib.c4ref_get_translate(nt2d, caa)
        params["nt2d"] = bytes(nt2d).hex()
        params["codon_aa"] = bytes(caa).hex()
        a, b, c, d, e = (C.c_int() for _ in range(5))
        lib.c4ref_get_intron_params(C.byref(a), C.byref(b), C.byref(c))
        params.update(min_intron=a.value, max_intron=b.value, intron_open=c.value)
        lib.c4ref_get_gap_params(C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e))
        params.update(gap_open=a.value, gap_extend=b.value, codon_gap_open=c.value,
                      codon_gap_extend=d.value, frameshift=e.value)
        with open(os.path.join(HERE, "scoring.json"), "w") as f:
            json.dump(params, f)

        kat_p2g = read_kat_strings("/root/reference/src/model/protein2genome.test.c")
        kat_c2c = read_kat_strings("/root/reference/src/model/coding2coding.test.c")

        # ---- models + cases -------------------------------------------------
        for fname, mname, qp, tp in MODELS:
            m = refdrv.RefModel(lib, mname, qp, tp, compiled=True)
            with open(os.path.join(HERE, "models", fname + ".txt"), "w") as f:
                f.write(m.dump())
            inputs = []
            if fname.startswith("affine") or fname.startswith("ungapped"):
                if qp:
                    inputs.append(("kat", KAT_AFFINE_Q, KAT_AFFINE_T))
                    for seed in range(12):
                        ql = [5, 17, 40, 64, 90, 130][seed % 6]
                        tl = [7, 33, 64, 50, 200, 95][seed % 6]
                        inputs.append(("rand%d" % seed,) + helpers.protein_pair(1000 + seed, ql, tl))
                else:
                    inputs.append(("tiny", "ACGT", "ACGT"))
                    inputs.append(("one", "A", "A"))
                    inputs.append(("mismatch", "AAAA", "CCCC"))
                    inputs.append(("ambig", "ACGTNNACGTRYACGT", "ACGTACNNGTACGTKM"))
                    for seed in range(16):
                        ql = [8, 31, 32, 33, 64, 100, 257, 300][seed % 8]
                        tl = [12, 64, 31, 200, 65, 99, 300, 1000][seed % 8]
                        inputs.append(("rand%d" % seed,) + helpers.dna_pair(2000 + seed, ql, tl))
                    # unrelated sequences (weak local hits, many ties)
                    rng = random.Random(77)
                    for k in range(4):
                        inputs.append(("noise%d" % k, helpers.rand_dna(rng, 50 + 30 * k),
                                       helpers.rand_dna(rng, 120 + 50 * k)))
                    # low-complexity: maximal tie-breaking stress
                    inputs.append(("lowc0", "ACACACACACACACACACAC", "CACACACACATTACACACACACACA"))
                    inputs.append(("lowc1", "AAAAAAAAAAAAAAAA", "AAAAAAAATAAAAAAAAAAAAAAA"))
                    inputs.append(("lowc2", "GGGGGGGGCCCCCCCC", "GGGGCCCCGGGGCCCCGGGGCCCC"))
            elif fname == "est2genome":
                inputs.append(("kat", KAT_E2G_Q, KAT_E2G_T))
                for seed in range(6):
                    inputs.append(("gene%d" % seed,) + est_genome_pair(3000 + seed,
                                                                       n_exons=2 + seed % 3,
                                                                       exon_len=40 + 10 * seed,
                                                                       intron_len=60 + 25 * seed))
                rq, rt = est_genome_pair(3100)
                comp = str.maketrans("ACGT", "TGCA")
                # cDNA from the opposite strand: exercises the *_reverse states
                inputs.append(("revgene", rq, rt.translate(comp)[::-1]))
            elif fname == "protein2genome":
                inputs.append(("kat", kat_p2g["query_seq"], kat_p2g["target_seq"]))
            elif fname == "coding2coding":
                inputs.append(("kat", kat_c2c["query_seq"], kat_c2c["target_seq"]))
                for seed in range(4):
                    q, t = helpers.dna_pair(4000 + seed, 60 + 9 * seed, 90 + 12 * seed, rate=0.08)
                    inputs.append(("rand%d" % seed, q, t))
            cases = []
            splice = {}
            for cname, q, t in inputs:
                p = m.pair(q, t, "qy", "tg")
                score = p.score()
                path = p.path(max_ops=1 << 14)
                case = {"name": cname, "q": q, "t": t, "score": score, "path": path}
                # sub-optimal series with blocking (a10), affine local DNA only
                if fname == "affine_local_dna" and cname in ("rand5", "rand7", "lowc0"):
                    p2 = m.pair(q, t, "qy", "tg")
                    series = []
                    for _ in range(3):
                        r = p2.path(use_subopt=True, add_to_subopt=True, max_ops=1 << 14)
                        if r is None:
                            break
                        series.append(r)
                    case["subopt_series"] = series
                    p2.close()
                cases.append(case)
                p.close()
                if fname in ("est2genome", "protein2genome"):
                    for ty in range(4):
                        arr = (C.c_int * len(t))()
                        lib.c4ref_splice_array(ty, t.encode(), len(t), arr)
                        splice["%s_%d" % (cname, ty)] = np.array(arr, dtype=np.int32)
            with open(os.path.join(HERE, "cases_" + fname + ".json"), "w") as f:
                json.dump(cases, f, indent=0)
            if splice:
                np.savez_compressed(os.path.join(HERE, "splice_" + fname + ".npz"), **splice)
            print(fname, len(cases), "cases")
            m.close()
        return 0

    refdrv.session(run)


if __name__ == "__main__":
    main()
