#!/usr/bin/env python
"""Goldens for the CLI batch hook at the metric shape (1 kbp x 100 kbp): the UNMODIFIED
reference binary (oracle/_ref/exonerate_c, compiled models) on the seeded FASTA files of
tests/cli_workload.py.  Only the outputs are committed; the GPU tier regenerates the same
FASTA files and requires byte-identical stdout from integration/_build/exonerate_b200.

Runs only in the build container (minutes of CPU: 4.5 s per affine lattice, 10 s per
est2genome lattice).  Usage: python tests/golden/make_cli_batch_golden.py [name ...]
"""
import concurrent.futures
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cli_workload  # noqa: E402

OUT = os.path.join(HERE, "cli_batch")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "exonerate_c")


def make(name):
    kind, nq, nt, flags = cli_workload.BATCH_COMMANDS[name][:4]
    with tempfile.TemporaryDirectory() as d:
        q, t = cli_workload.write_workload(d, kind, nq, nt, *cli_workload.BATCH_COMMANDS[name][4:])
        out = subprocess.run([REF_BIN, q, t] + flags + cli_workload.COMMON, capture_output=True, text=True,
                             check=True).stdout
    with open(os.path.join(OUT, name + ".out"), "w") as f:
        f.write(out)
    return name, len(out.splitlines())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or sorted(cli_workload.BATCH_COMMANDS)
    with concurrent.futures.ThreadPoolExecutor(max_workers=4) as ex:
        for name, lines in ex.map(make, names):
            print(name, lines, "lines")
