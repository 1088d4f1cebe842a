"""GPU tier: the exonerate CLI itself, bit for bit.

integration/_build/exonerate_b200 is the UNMODIFIED reference (hub, models,
alignment printers, FASTA I/O ...) with OUR viterbi.o replacement linked in place
of the reference's, so every lattice fill and traceback of an
`--exhaustive` run happens in libc4b200.so.  Its stdout must equal, byte for
byte, what the reference's own binary (compiled models) printed for the same
command lines: scores, coordinates, strands, cigar and vulgar strings, --ryo
percent identity, sub-optimal series (--subopt yes), reverse-complement passes.

The bsdp_* command lines run the heuristic path instead (--exhaustive no
--gappedextension no): HSP seeding and SAR stay the reference's host code, and
every DP they ask for -- the bound fills of Heuristic_create (END-cell table),
terminal and join fills on derived models with SubOpt blocking, the span models of
est2genome / protein2genome / coding2coding with their cell callbacks -- is ours.

The binary exists only where the reference was present at build time; it
travels to the GPU box with the snapshot.  Skipped (not failed) if absent."""
import json
import os
import subprocess

import pytest

import helpers

pytestmark = pytest.mark.gpu

CLI = os.path.join(helpers.GOLDEN, "cli")
BIN = os.path.join(helpers.ROOT, "integration", "_build", "exonerate_b200")
COMMANDS = json.load(open(os.path.join(CLI, "commands.json")))


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration binary not built (needs the reference sources)")
@pytest.mark.parametrize("name", sorted(COMMANDS))
def test_cli_output_identical_to_reference(name):
    want = open(os.path.join(CLI, name + ".out")).read()
    got = subprocess.run([BIN] + COMMANDS[name], cwd=CLI, capture_output=True, text=True, timeout=600)
    assert got.returncode == 0, got.stderr[-2000:]
    assert got.stdout == want
    assert len(want.splitlines()) >= 2


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration binary not built (needs the reference sources)")
@pytest.mark.parametrize("name", ["bsdp_affine_local_dna", "bsdp_protein2genome", "ungapped_dna"])
def test_cli_heuristic_runs_extend_hsps_on_the_device(name):
    """the hspset binding (integration/hspset_b200.c) really is on the path: the seeds of
    DNA2DNA / PROTEIN2DNA comparisons are extended by c4b_hsp_extend_batch"""
    import re
    env = dict(os.environ, EXONERATE_B200_STATS="1")
    got = subprocess.run([BIN] + COMMANDS[name], cwd=CLI, capture_output=True, text=True, timeout=600, env=env)
    assert got.returncode == 0, got.stderr[-2000:]
    m = re.search(r"hsp batches (\d+), seeds extended on the device (\d+), seeds passed to the reference (\d+)",
                  got.stderr)
    assert m, got.stderr[-500:]
    assert int(m.group(1)) >= 1 and int(m.group(2)) >= 1 and int(m.group(3)) == 0


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration binary not built (needs the reference sources)")
@pytest.mark.parametrize("name", ["bsdp_protein2genome", "bsdp_est2genome"])
def test_cli_heuristic_runs_on_specialised_kernels(name, tmp_path):
    """BSDP through the run-time SPECIALISED kernels: every derived model (terminal, join, span
    with its START / END cell tables, SubOpt blocking) is compiled for sm_100a on first use and
    kept in a disk cache; the second process must find all of them there, the third one finds one
    entry corrupted and replaces it.  Output byte-identical to the reference every time."""
    want = open(os.path.join(CLI, name + ".out")).read()
    cache = tmp_path / "jit"
    cache.mkdir()
    env = dict(os.environ, C4B_JIT_CACHE_DIR=str(cache), EXONERATE_B200_STATS="1")
    for attempt in range(3):
        got = subprocess.run([BIN] + COMMANDS[name], cwd=CLI, capture_output=True, text=True, timeout=900, env=env)
        assert got.returncode == 0, got.stderr[-2000:]
        assert got.stdout == want
        assert "interpreter kernel" not in got.stderr  # no specialisation failed
        if attempt == 0:
            cubins = sorted(os.listdir(cache))
            assert len(cubins) >= 3
        else:
            assert sorted(os.listdir(cache)) == cubins
        if attempt == 1:  # a cache entry that no longer loads is recompiled, not trusted
            with open(cache / cubins[0], "wb") as f:
                f.write(b"not a cubin")
    assert os.path.getsize(cache / cubins[0]) > 1000


# ---- the batch hook (integration/gam_b200.c) at the metric shape -------------------------
import cli_workload  # noqa: E402

BATCH_GOLDEN = os.path.join(helpers.GOLDEN, "cli_batch")


def _run_batch_cli(tmp_path, name, extra_env=None):
    kind, nq, nt, flags = cli_workload.BATCH_COMMANDS[name][:4]
    q, t = cli_workload.write_workload(str(tmp_path), kind, nq, nt, *cli_workload.BATCH_COMMANDS[name][4:])
    env = dict(os.environ, EXONERATE_B200_STATS="1", **(extra_env or {}))
    return subprocess.run([BIN, q, t] + flags + cli_workload.COMMON, capture_output=True, text=True, timeout=900,
                          env=env)


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration binary not built (needs the reference sources)")
@pytest.mark.parametrize("name", sorted(cli_workload.BATCH_COMMANDS))
def test_cli_batch_hook_metric_shape(name, tmp_path):
    """`exonerate --exhaustive` on 1 kbp x 100 kbp FASTA input through the batch hook: the queued
    pairs are answered by batched device rounds and replayed through the reference's own
    GAM_Result_exhaustive_create; stdout byte-identical to the reference binary (compiled models)."""
    import re
    want = open(os.path.join(BATCH_GOLDEN, name + ".out")).read()
    got = _run_batch_cli(tmp_path, name)
    assert got.returncode == 0, got.stderr[-2000:]
    assert got.stdout == want
    assert len(want.splitlines()) >= 4
    m = re.search(r"batch hook: (\d+) pair\(s\) in (\d+) flush\(es\), (\d+) device round", got.stderr)
    assert m, got.stderr[-800:]
    kind, nq, nt, flags = cli_workload.BATCH_COMMANDS[name][:4]
    strands = 1 if "--revcomp" in flags else 2   # --revcomp defaults to yes: every query twice
    assert int(m.group(1)) == nq * nt * strands and int(m.group(2)) == 1
    m = re.search(r"answered from the batch prefetch (\d+) \(prefetched but not usable (\d+)\)", got.stderr)
    assert m and int(m.group(1)) >= nq * nt * strands and int(m.group(2)) == 0, got.stderr[-800:]
    m = re.search(r"Viterbi_calculate calls (\d+) .* answered from the batch prefetch (\d+)", got.stderr)
    assert m, got.stderr[-800:]
    if "--bestn" not in flags and "--subopt" in flags:   # every Viterbi_calculate of the run came from a batch
        assert m.group(1) == m.group(2), got.stderr[-800:]
    elif "--bestn" not in flags:   # sub-optimal series deeper than the 16 prefetched rounds run on synchronously
        assert int(m.group(2)) >= 0.95 * int(m.group(1)), got.stderr[-800:]


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration binary not built (needs the reference sources)")
def test_cli_batch_hook_small_queue_and_off(tmp_path):
    """a queue bound of 5 pairs (several flushes) and batching switched off print the same bytes"""
    name = "metric_affine_local"
    want = open(os.path.join(BATCH_GOLDEN, name + ".out")).read()
    got = _run_batch_cli(tmp_path, name, {"EXONERATE_B200_BATCH_PAIRS": "5"})
    assert got.returncode == 0 and got.stdout == want, got.stderr[-2000:]
    assert "in 5 flush(es)" in got.stderr
    got = _run_batch_cli(tmp_path, name, {"EXONERATE_B200_BATCH": "0"})
    assert got.returncode == 0 and got.stdout == want, got.stderr[-2000:]
    # the flushes sharded over a device group inside the C library (two engines on GPU 0 here)
    got = _run_batch_cli(tmp_path, name, {"EXONERATE_B200_DEVICES": "0,0"})
    assert got.returncode == 0 and got.stdout == want, got.stderr[-2000:]
