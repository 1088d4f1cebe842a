#!/usr/bin/env python
"""bench.py -- GCUPS of the C4 affine:local hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference --gpus N --steps K ...   reference CPU path

One "step" = Optimal_find_path over one batch of synthetic 1 kbp x 100 kbp DNA
pairs (score + alignment region + operation list for every pair).  Prints ONE
JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "GCUPS affine:local find_path (score+region+ops, bit-exact), 1 kbp x 100 kbp DNA batch"
B_ALG = 20  # algorithmic bytes per lattice cell, SURVEY.md 8d: 4 B x 5 states x C=1

# The default run is BASELINE.json's metric configuration (affine:local).  --model
# est2genome runs north_star's second target workload (configs[2]) the same way.
WORKLOADS = {
    "affine:local": {
        "metric": METRIC, "b_alg": B_ALG, "pairs": 10000, "cpu_pairs": 2,
        "workload": "affine:local --exhaustive, %d bp x %d bp DNA pairs",
        "kernel": "affine_fill16u (score pass, two lattices per warp in 16-bit halves) + affine_fill16tb "
                  "(banded traceback pass, tagged 16-bit halves) + walk",
        "traffic_key": "affine_fill16u_kernel<32>",
        "note": "B_alg=20 B/cell (SURVEY 8d, reference row layout); the kernel keeps rows in registers and "
                "writes 0.5 B/cell only inside the traceback band, so frac>1 is expected; see DESIGN.md. peak ",
    },
    "est2genome": {
        "metric": "GCUPS est2genome find_path (score+region+ops, bit-exact), 1 kbp cDNA x 100 kbp genomic batch",
        "b_alg": 80, "pairs": 1000, "cpu_pairs": 1,
        "workload": "est2genome --exhaustive, %d bp cDNA x %d bp genomic pairs (5 exons, GT..AG introns)",
        "kernel": "e2g_fill16 (both strands per register, score pass + column checkpoints) + window refills "
                  "with 15-bit traceback records under the path + walk with intron jumps",
        "traffic_key": "e2g_fill16_kernel<2>",
        "note": "B_alg=80 B/cell (SURVEY 8d: 4 B x 10 states x C=2, reference row layout); the kernel keeps "
                "rows in registers, writes 28 B/row checkpoints every 1024 columns and 2 B/cell records only "
                "inside the refilled windows, so frac>1 is expected; see DESIGN.md. peak ",
    },
    # the table-driven path (any closed model), specialised per model at run time
    "protein2genome": {
        "metric": "GCUPS protein2genome find_path (score+region+ops, bit-exact), 450 aa x 20 kbp genomic batch",
        "b_alg": 104, "pairs": 592, "cpu_pairs": 1, "qlen": 450, "tlen": 20000, "query_is_protein": True,
        "workload": "protein2genome --exhaustive, %d aa protein x %d bp genomic pairs (4 exons, GT..AG introns)",
        "kernel": "c4b_jit_sys (closed model compiled for sm_100a at run time as a systolic kernel: lattice in "
                  "registers, lanes skewed by a column, one PATH pass with rank-bit records) + walk",
        "traffic_key": "c4b_jit_sys<protein2genome>",
        "note": "B_alg=104 B/cell (SURVEY 8d: 4 B x 13 states x C=2, reference row layout); the kernel keeps the "
                "lattice in registers and writes 4 B/cell of rank-bit traceback records, so the HBM roofline does "
                "not bind it (ALU-pipe bound, profiles/r02_kernels.md). peak ",
    },
}


# ----------------------------------------------------------------------------
# synthetic workload: random target, mutated query planted in it
# ----------------------------------------------------------------------------
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_batch(seed, n, qlen, tlen, rate=0.15):
    rng = np.random.default_rng(seed)
    queries = ACGT[rng.integers(0, 4, size=(n, qlen), dtype=np.uint8)]
    targets = ACGT[rng.integers(0, 4, size=(n, tlen), dtype=np.uint8)]
    for k in range(n):
        q = queries[k]
        u = rng.random(qlen)
        kind = np.zeros(qlen, dtype=np.int8)          # 0 keep
        kind[u < rate] = 1                            # 1 delete
        kind[u < 2 * rate / 3] = 2                    # 2 insert before
        kind[u < rate / 3] = 3                        # 3 substitute
        counts = np.where(kind == 1, 0, np.where(kind == 2, 2, 1))
        core = np.repeat(q, counts)
        pos = np.cumsum(counts) - counts
        rnd = ACGT[rng.integers(0, 4, size=qlen)]
        sel = (kind == 2) | (kind == 3)
        core[pos[sel]] = rnd[sel]
        core = core[:tlen]
        off = int(rng.integers(0, tlen - len(core) + 1))
        targets[k, off:off + len(core)] = core
    return queries, targets


def make_batch_e2g(seed, n, qlen, tlen, n_exons=5, rate=0.02):
    """cDNA of n_exons exons; genomic = random sequence with
    the (lightly mutated) exons planted in order, separated by GT...AG introns."""
    rng = np.random.default_rng(seed)
    exon = qlen // n_exons
    qlen = exon * n_exons
    intron = max(40, (tlen * 3 // 4 - qlen) // max(1, n_exons - 1))
    body_len = qlen + (n_exons - 1) * (intron + 4)
    queries = ACGT[rng.integers(0, 4, size=(n, qlen), dtype=np.uint8)]
    targets = ACGT[rng.integers(0, 4, size=(n, max(tlen, body_len)), dtype=np.uint8)]
    for k in range(n):
        off = int(rng.integers(0, targets.shape[1] - body_len + 1))
        pos = off
        for e in range(n_exons):
            ex = queries[k, e * exon:(e + 1) * exon].copy()
            sub = rng.random(exon) < rate
            ex[sub] = ACGT[rng.integers(0, 4, size=int(sub.sum()))]
            targets[k, pos:pos + exon] = ex
            pos += exon
            if e + 1 < n_exons:
                targets[k, pos:pos + 2] = np.frombuffer(b"GT", dtype=np.uint8)
                targets[k, pos + 2 + intron:pos + 4 + intron] = np.frombuffer(b"AG", dtype=np.uint8)
                pos += intron + 4
    return queries, targets


CODON = {"A": "GCT", "R": "CGT", "N": "AAC", "D": "GAC", "C": "TGC", "Q": "CAG", "E": "GAG", "G": "GGT",
         "H": "CAC", "I": "ATC", "L": "CTG", "K": "AAG", "M": "ATG", "F": "TTC", "P": "CCG", "S": "TCT",
         "T": "ACC", "W": "TGG", "Y": "TAC", "V": "GTT"}


def make_batch_p2g(seed, n, aa, tlen, n_exons=4, rate=0.03):
    """SURVEY.md 8d config 4 shape: random proteins; genomic = random sequence with the
    back-translated protein planted as n_exons exons (cut at codon boundaries), GT...AG introns."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer("".join(CODON).encode(), dtype=np.uint8)
    table = np.zeros((256, 3), dtype=np.uint8)
    for a, c in CODON.items():
        table[ord(a)] = np.frombuffer(c.encode(), dtype=np.uint8)
    queries = letters[rng.integers(0, 20, size=(n, aa))]
    targets = ACGT[rng.integers(0, 4, size=(n, tlen), dtype=np.uint8)]
    intron = max(60, (tlen // 2 - 3 * aa) // max(1, n_exons - 1))
    body_len = 3 * aa + (n_exons - 1) * (intron + 4)
    assert body_len <= tlen
    for k in range(n):
        cds = table[queries[k]].reshape(-1).copy()
        sub = rng.random(cds.size) < rate
        cds[sub] = ACGT[rng.integers(0, 4, size=int(sub.sum()))]
        cuts = [0] + sorted(3 * int(c) for c in rng.choice(np.arange(10, aa - 10), n_exons - 1, replace=False)) + [3 * aa]
        pos = int(rng.integers(0, tlen - body_len + 1))
        for e in range(n_exons):
            ex = cds[cuts[e]:cuts[e + 1]]
            targets[k, pos:pos + ex.size] = ex
            pos += ex.size
            if e + 1 < n_exons:
                targets[k, pos:pos + 2] = np.frombuffer(b"GT", dtype=np.uint8)
                targets[k, pos + 2 + intron:pos + 4 + intron] = np.frombuffer(b"AG", dtype=np.uint8)
                pos += intron + 4
    return queries, targets


GENERATORS = {"affine:local": make_batch, "est2genome": make_batch_e2g, "protein2genome": make_batch_p2g}


def clocks_sampler(stop, out, index):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                              "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return
    def reader():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop.wait()
    p.terminate()
    t.join(timeout=2)


def summarise_clocks(lines):
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 6:
            continue
        try:
            sm.append(float(f[0]))
            mx = max(mx, float(f[1]))
        except ValueError:
            continue
        for nm, v in zip(names, f[2:6]):
            if v.lower().startswith("active"):
                reasons.add(nm)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
            "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic(cells, key):
    """ncu-measured DRAM bytes of the dominant kernel, scaled to this launch."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        rec = json.load(open(path))[key]
        return (rec["dram_bytes_read"] + rec["dram_bytes_write"]) / rec["cells"] * cells
    except (OSError, KeyError, ValueError):
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------
# CPU arms: the reference's own implementation of the path on host cores
# ----------------------------------------------------------------------------
def _cpu_worker(kind, conn, model_name="affine:local"):
    """One host process = one single-threaded reference instance.  Receives
    lists of (query, target) strings, answers (seconds, scores)."""
    def serve(align):
        while True:
            task = conn.recv()
            if task is None:
                return 0
            t0 = time.perf_counter()
            scores = [align(q, t) for q, t in task]
            conn.send((time.perf_counter() - t0, scores))

    if kind == "reference":
        # the UNMODIFIED reference (compiled models), one long session per process
        from oracle import refdrv

        def fn(lib):
            m = refdrv.RefModel(lib, model_name, int(WORKLOADS[model_name].get("query_is_protein", False)), 0,
                                compiled=True)

            def align(q, t):
                p = m.pair(q, t)
                r = p.path(max_ops=1 << 15)
                p.close()
                return (r["score"], r["region"], r["ops"])
            return serve(align)
        refdrv.session(fn)
    else:
        # the oracle port (only where the reference binary did not travel)
        import helpers
        params = helpers.load_params()
        scoring = helpers.load_scoring(params)
        model, _ = helpers.load_model("affine_local_dna" if model_name == "affine:local" else model_name, params)

        def align(q, t):
            sp = None
            if model_name in ("est2genome", "protein2genome"):
                from exonerate_b200.models import splice_arrays
                sp = splice_arrays(t)
            r = helpers.oracle_find_path(model, scoring, helpers.PairBuf(q, t, splice=sp),
                                         region_threshold_cells=0, max_ops=len(q) + len(t) + 8)
            return (r["score"], r["region"], r["ops"])
        serve(align)


def cpu_arm_available():
    from oracle import refdrv
    return "reference" if refdrv.available() else "port"


class CpuPool:
    """`procs` persistent reference processes (the reference has no threading)."""

    def __init__(self, kind, procs, model_name="affine:local"):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        self.kind, self.workers = kind, []
        for _ in range(procs):
            parent, child = ctx.Pipe()
            pr = ctx.Process(target=_cpu_worker, args=(kind, child, model_name), daemon=True)
            pr.start()
            self.workers.append((pr, parent))

    def run(self, queries, targets, full=False):
        """Wall time for all pairs, dealt round-robin to the processes.  Answers are scores, or
        (score, region, ops) with full=True."""
        n, w = len(queries), len(self.workers)
        tasks = [[] for _ in range(w)]
        for k in range(n):
            tasks[k % w].append((bytes(queries[k]).decode(), bytes(targets[k]).decode()))
        t0 = time.perf_counter()
        used = [(c, t) for (_, c), t in zip(self.workers, tasks) if t]
        for c, t in used:
            c.send(t)
        res = [c.recv() for c, _ in used]
        wall = time.perf_counter() - t0
        scores = [None] * n
        for wi, (_, sc) in enumerate(res):
            for j, s in enumerate(sc):
                scores[wi + j * w] = s if full else s[0]
        return wall, scores

    def close(self):
        for pr, c in self.workers:
            try:
                c.send(None)
            except (BrokenPipeError, OSError):
                pass
        for pr, _ in self.workers:
            pr.join(timeout=5)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # only rank 0 runs the CPU arm under torchrun
    kind = cpu_arm_available()
    procs = max(1, min(os.cpu_count() or 1, 64))
    W = WORKLOADS[args.model]
    n = procs  # one pair per host process per step: a bounded sample of the workload
    gen = GENERATORS[args.model]
    queries, targets = gen(1000, n, args.qlen, args.tlen)  # same generator/seed as rank 0 of our arm
    cells = n * queries.shape[1] * targets.shape[1]
    pool = CpuPool(kind, procs, args.model)
    for _ in range(args.warmup):
        pool.run(queries, targets)
    times = [pool.run(queries, targets)[0] for _ in range(args.steps)]
    pool.close()
    t = float(np.mean(times))
    gcups = cells / t / 1e9
    line = {"impl": "reference", "metric": W["metric"], "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": W["workload"] % (args.qlen, args.tlen), "pairs_per_step": n},
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": procs, "kind": kind,
                             "sample": "%d pairs of %d x %d per step, one per host process" % (n, args.qlen, args.tlen)},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
# ALU-issue roofline of the dominant kernels: warp-instructions per lattice cell from ncu
# (smsp__inst_executed.sum / cells of the same launch; profiles/traffic.json) against the issue
# rate of the chip, SMs x 4 schedulers x SM clock -- the bound that actually limits kernels
# that keep the lattice in registers (pipe_alu 79 %, issue 76 %).
def alu_roofline(cells, ms, key, sm_count, sm_mhz):
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
        ipc = rec["warp_inst"] / rec["cells"]
    except (OSError, KeyError, ValueError):
        return None
    peak = sm_count * 4 * (sm_mhz or 1965.0) * 1e6 / 1e9
    achieved = ipc * cells / (ms * 1e-3) / 1e9
    return {"bound": "alu-issue", "achieved": achieved, "peak": peak, "unit": "G warp-inst/s",
            "frac": achieved / peak, "warp_inst_per_cell": ipc,
            "source": "ncu smsp__inst_executed.sum / cells of %s (%s); pipe_alu %.1f %%, issue_active %.1f %% in "
                      "that capture; peak = %d SMs x 4 schedulers x %.0f MHz" % (
                          key, rec.get("profile", "profiles/"), rec.get("pipe_alu_pct", 0),
                          rec.get("issue_active_pct", 0), sm_count, sm_mhz or 1965.0)}


class Job:
    """rank / world / timing plumbing shared by every workload of one bench process"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the C4 fill has no CPU fallback)")
        torch.cuda.set_device(self.local)
        # host-side staging threads of the engine: the ranks of one box share its cores
        os.environ.setdefault("C4B_HOST_THREADS", str(max(1, min(16, (os.cpu_count() or 16) // self.world))))
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.sm_count = torch.cuda.get_device_properties(self.local).multi_processor_count

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t]


def run_workload(job, eng, model_name, n, qlen, tlen, steps, warmup, sample_clocks=False, want_e2e=True):
    """One workload on this rank's shard of n pairs: device-resident steps (CUDA events on the
    launching stream) and end-to-end steps through c4b_find_path_batch with host buffers.
    Returns a dict of per-job numbers (max over ranks for times, sum for cells)."""
    torch = job.torch
    import helpers
    from exonerate_b200 import Batch, Optimal, PairSet, abi
    from exonerate_b200.models import host_model
    from exonerate_b200.sharding import gather_records
    W = WORKLOADS[model_name]
    params = helpers.load_params()
    scoring = helpers.load_scoring(params)   # the reference's Submat tables (tests/golden/scoring.json)
    model, _ = host_model(model_name, query_is_protein=W.get("query_is_protein", False))  # closed by csrc/host
    queries, targets = GENERATORS[model_name](1000 + job.rank, n, qlen, tlen)
    splice = None
    if model_name != "affine:local":
        from exonerate_b200.models import splice_arrays
        splice = [splice_arrays(targets[k]) for k in range(n)]   # host C splice predictor (csrc/host/splice.c)
    if model_name == "protein2genome":
        os.environ.setdefault("C4B_GENERIC_JIT", "1")  # the batch is below the auto-specialise size
    # the step's inputs live in pinned host memory (bench contract): the engine DMAs straight from it
    pin = [torch.empty(a.shape, dtype=torch.uint8, pin_memory=True) for a in (queries, targets)]
    for t_, a in zip(pin, (queries, targets)):
        t_.numpy()[...] = a
    queries, targets = pin[0].numpy(), pin[1].numpy()
    pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)], splice=splice, pinned=True)
    opt = Optimal(eng, model, scoring)
    shards = [np.arange(r * n, (r + 1) * n) for r in range(job.world)]

    # ---- device-resident arm: inputs staged in HBM before the timed region ----
    batch = Batch(eng, model, scoring, pairs, want_path=True)
    batch.run()
    dev_results = torch.as_tensor(batch.device_results(), device="cuda")  # c4b_result[n] in HBM, zero-copy

    def step_resident():
        batch.run()
        if job.world > 1:  # per-pair result records gathered over NCCL/NVLink (north_star)
            gather_records(dev_results, shards, job.rank, job.world)
    for _ in range(warmup):
        step_resident()
    job.barrier()
    stop, lines, sampler = threading.Event(), [], None
    if sample_clocks:
        sampler = threading.Thread(target=clocks_sampler, args=(stop, lines, job.local), daemon=True)
        sampler.start()
    launches0 = eng.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fill_ms = []
    e0.record()
    for _ in range(steps):
        step_resident()
        fill_ms.append(batch.last_fill_ms())   # fill-kernel events of THIS step (syncs the stream)
    e1.record()
    job.barrier()
    launches = eng.kernel_launches() - launches0
    dev_ms = e0.elapsed_time(e1) / steps
    stop.set()
    if sampler:
        sampler.join(timeout=3)
    need = batch.ops_needed()
    results, ops = batch.fetch(ops_capacity=need)
    kernel_name, route = batch.kernel_name, batch.description
    batch.close()
    out = {"n": n, "cells": pairs.cells, "dev_ms": dev_ms, "fill_ms": float(np.mean(fill_ms)), "launches": launches,
           "clock_lines": lines, "results": results, "ops": ops, "queries": queries, "targets": targets,
           "pairs": pairs, "n_ops_total": int(need), "kernel_name": kernel_name, "route": route, "e2e_ms": None, "_pinned": pin}

    # ---- end-to-end arm: host buffers in, host results out, every step ----
    if want_e2e:
        hout = ((abi.Result * n)(), np.empty(2 * max(int(need), 1) + 2, dtype=np.int32))  # reused host buffers
        host_scores = torch.zeros((n, 10), dtype=torch.int32, device="cuda")
        for _ in range(min(warmup, 2)):
            opt.find_path_raw(pairs, out=hout)
        job.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            got, _ = opt.find_path_raw(pairs, out=hout)
            if job.world > 1:
                host_scores.copy_(torch.from_numpy(np.frombuffer(got, dtype=np.int32).reshape(n, 10)))
                gather_records(host_scores, shards, job.rank, job.world)
        job.barrier()
        out["e2e_ms"] = (time.perf_counter() - t0) / steps * 1e3
        for k in range(n):   # the two arms agree record for record
            a, b_ = got[k], results[k]
            assert (a.score, a.query_start, a.target_start, a.query_end, a.target_end, a.n_ops) == \
                   (b_.score, b_.query_start, b_.target_start, b_.query_end, b_.target_end, b_.n_ops), k
    dev_ms, e2e_ms = job.max_over_ranks([out["dev_ms"], out["e2e_ms"] or 0.0])
    (total_cells,) = job.sum_over_ranks([float(pairs.cells)])
    out.update(dev_ms=dev_ms, e2e_ms=e2e_ms if want_e2e else None, total_cells=total_cells,
               value=total_cells / (dev_ms * 1e-3) / 1e9,
               e2e_value=(total_cells / (e2e_ms * 1e-3) / 1e9) if want_e2e else None)
    return out


def check_against_reference(model_name, w, k):
    """the first k pairs of the batch on the reference's own CPU implementation (one host
    core): score, alignment region AND operation list must equal the GPU's"""
    kind = cpu_arm_available()
    pool = CpuPool(kind, 1, model_name)
    wall, cpu = pool.run(w["queries"][:k], w["targets"][:k], full=True)
    pool.close()
    for i in range(k):
        r = w["results"][i]
        o = int(r.ops_offset)
        gpu = (r.score, [r.query_start, r.target_start, r.query_end - r.query_start, r.target_end - r.target_start],
               [(int(w["ops"][2 * (o + j)]), int(w["ops"][2 * (o + j) + 1])) for j in range(r.n_ops)])
        assert cpu[i][0] == gpu[0], "CPU baseline score differs for pair %d: %r vs %r" % (i, cpu[i][0], gpu[0])
        if cpu[i][1] is not None:   # (the oracle port returns regions and ops too)
            assert list(cpu[i][1]) == gpu[1], "CPU baseline region differs for pair %d" % i
            assert [tuple(x) for x in cpu[i][2]] == gpu[2], "CPU baseline ops differ for pair %d" % i
    q, t = w["queries"].shape[1], w["targets"].shape[1]
    return {"value": k * q * t / wall / 1e9, "unit": "GCUPS", "cores": 1, "kind": kind,
            "sample": "first %d pair(s) of the batch, single-threaded, %.1f s; score, region and ops equal the GPU's"
                      % (k, wall)}


def cli_leg(n_queries=100, n_targets=100, check_queries=1, check_targets=3):
    """The shipped binary: integration/_build/exonerate_b200 (the unmodified reference with our
    viterbi.o + the batch hook gam_b200.o) on a FASTA of n_queries x n_targets 1 kbp x 100 kbp
    sequences, wall clock of the whole process; a subsample is run through the reference's own
    binary (oracle/_ref/exonerate_c, compiled models) and must print the same bytes."""
    import re
    import tempfile
    import cli_workload
    exe = os.path.join(ROOT, "integration", "_build", "exonerate_b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "exonerate_c")
    if not os.path.exists(exe):
        return {"unavailable": "integration/_build/exonerate_b200 not built (needs the reference sources)"}
    flags = ["--model", "affine:local", "--exhaustive", "yes", "--subopt", "no", "--revcomp", "no", "--score", "0"]
    with tempfile.TemporaryDirectory() as d:
        rng = np.random.default_rng(77)
        qs, ts = make_batch(77, max(n_queries, n_targets), 1000, 100000)
        qrec = [("q%d" % k, bytes(qs[k]).decode()) for k in range(n_queries)]
        trec = [("t%d" % k, bytes(ts[k]).decode()) for k in range(n_targets)]   # t_k holds a copy of q_k
        q, t = os.path.join(d, "q.fa"), os.path.join(d, "t.fa")
        cli_workload.write_fasta(q, qrec)
        cli_workload.write_fasta(t, trec)
        env = dict(os.environ, EXONERATE_B200_STATS="1")
        t0 = time.perf_counter()
        got = subprocess.run([exe, q, t] + flags + cli_workload.COMMON, capture_output=True, text=True, env=env)
        wall = time.perf_counter() - t0
        if got.returncode != 0:
            return {"error": got.stderr[-500:]}
        cells = n_queries * n_targets * 1000 * 100000
        res = {"binary": "integration/_build/exonerate_b200", "pairs": n_queries * n_targets,
               "flags": " ".join(flags), "wall_s": wall, "value": cells / wall / 1e9, "unit": "GCUPS",
               "alignments": got.stdout.count("vulgar:")}
        m = re.search(r"flatten ([\d.]+) s, splice arrays ([\d.]+) s, device ([\d.]+) s \(([\d.]+) GCUPS\), replay ([\d.]+) s",
                      got.stderr)
        if m:
            res.update(flatten_s=float(m.group(1)), device_s=float(m.group(3)), device_gcups=float(m.group(4)),
                       replay_s=float(m.group(5)),
                       other_s=wall - float(m.group(1)) - float(m.group(3)) - float(m.group(5)),
                       note="other_s = the reference's own FASTA parsing (fastapipe re-reads every target per "
                            "query), process and CUDA start-up, printing")
        if os.path.exists(ref) and check_queries:
            sq, st = os.path.join(d, "sq.fa"), os.path.join(d, "st.fa")
            cli_workload.write_fasta(sq, qrec[:check_queries])
            cli_workload.write_fasta(st, trec[:check_targets])
            t0 = time.perf_counter()
            want = subprocess.run([ref, sq, st] + flags + cli_workload.COMMON, capture_output=True, text=True).stdout
            ref_wall = time.perf_counter() - t0
            mine = subprocess.run([exe, sq, st] + flags + cli_workload.COMMON, capture_output=True, text=True,
                                  env=env).stdout
            assert want == mine and want.count("vulgar:") >= 1, "CLI output differs from the reference binary"
            # ... and the same records inside the big run
            for line in want.splitlines():
                if line.startswith(("vulgar:", "cigar:")):
                    assert line in got.stdout, "subsample line missing from the batch run: " + line[:80]
            res["checked"] = "%d x %d pairs byte-identical to oracle/_ref/exonerate_c (%.1f s, %.3f GCUPS)" % (
                check_queries, check_targets, ref_wall, check_queries * check_targets * 1e8 / ref_wall / 1e9)
        return res


def ours(args):
    from exonerate_b200 import Engine
    job = Job()
    torch = job.torch
    rank, world = job.rank, job.world
    W = WORKLOADS[args.model]
    strong = args.scaling == "strong"
    total_pairs = args.pairs or W["pairs"]
    n = max(1, total_pairs // world) if strong else total_pairs

    eng = Engine(job.local)
    eng.lib.c4b_engine_set_stream(eng.h, torch.cuda.current_stream().cuda_stream)
    w = run_workload(job, eng, args.model, n, args.qlen, args.tlen, args.steps, args.warmup, sample_clocks=True)

    extra, strong_block = {}, {}
    if not args.only_main:
        # north_star's other workloads in the same driver-run line (fewer steps: they are context)
        for name in ("est2genome", "protein2genome"):
            if name == args.model:
                continue
            WW = WORKLOADS[name]
            x = run_workload(job, eng, name, WW["pairs"], WW.get("qlen", 1000), WW.get("tlen", 100000),
                             steps=3, warmup=2)
            extra[name] = {"metric": WW["metric"], "value": x["value"], "unit": "GCUPS", "ms_per_step": x["dev_ms"],
                           "e2e": x["e2e_value"], "pairs_per_gpu": WW["pairs"], "scaling": "weak",
                           "kernel": x["kernel_name"], "b_alg_bytes_per_cell": WW["b_alg"],
                           "hbm_roofline_gcups_per_gpu": measured_peak()[0] / WW["b_alg"]}
        # strong scaling: north_star's FIXED batches split over the ranks
        for name, tot in (("affine:local", 10000), ("est2genome", 1000)):
            WW = WORKLOADS[name]
            if world == 1 and name == args.model and n == tot and not strong:
                x = w   # at N=1 the fixed batch IS the main workload: same numbers, not measured twice
            elif world == 1 and name in extra and WW["pairs"] == tot:
                strong_block[name] = {"pairs_total": tot, "pairs_per_gpu": tot, "value": extra[name]["value"],
                                      "e2e": extra[name]["e2e"], "unit": "GCUPS",
                                      "ms_per_step": extra[name]["ms_per_step"]}
                continue
            else:
                x = run_workload(job, eng, name, max(1, tot // world), WW.get("qlen", 1000), WW.get("tlen", 100000),
                                 steps=3, warmup=2)
            strong_block[name] = {"pairs_total": (tot // world) * world, "pairs_per_gpu": tot // world,
                                  "value": x["value"], "e2e": x["e2e_value"], "unit": "GCUPS",
                                  "ms_per_step": x["dev_ms"]}

    int32_value = None
    if args.model == "affine:local" and not args.only_main:
        # the same batch on the int32 kernels: what a lattice outside the 16-bit bound gets
        # (an N in the query, a protein, global / bestfit / overlap scopes, Q > 6399, SubOpt)
        os.environ["C4B_AFFINE_PACK16"] = "0"
        os.environ["C4B_AFFINE_TB16"] = "0"
        x = run_workload(job, eng, args.model, n, args.qlen, args.tlen, steps=3, warmup=2, want_e2e=False)
        del os.environ["C4B_AFFINE_PACK16"], os.environ["C4B_AFFINE_TB16"]
        int32_value = x["value"]
        assert all(x["results"][k].score == w["results"][k].score and x["results"][k].n_ops == w["results"][k].n_ops
                   for k in range(n)), "int32 and packed kernels disagree"

    if rank == 0:
        peak, peak_src = measured_peak()
        clocks = summarise_clocks(w["clock_lines"])
        achieved = w["cells"] * W["b_alg"] / (w["fill_ms"] * 1e-3) / 1e9
        q, t = w["queries"].shape[1], w["targets"].shape[1]
        line = {
            "metric": W["metric"], "value": w["value"], "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": w["dev_ms"], "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": W["workload"] % (q, t),
                       "pairs_per_gpu": n, "cells_per_step": w["total_cells"], "seed": 1000,
                       "l2": "inputs (%d MB per GPU) exceed the 126 MB L2" % (w["pairs"].h2d_bytes >> 20),
                       "kernel": W["kernel"],
                       "route": w["route"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(w["cells"], W["traffic_key"]),
                         "note": W["note"] + peak_src, "fill_kernel_ms": w["fill_ms"],
                         "fill_kernel_ms_is": "mean over the %d timed steps (CUDA events on the launching streams)"
                                              % args.steps},
            "roofline_alu": alu_roofline(w["cells"], w["fill_ms"], W["traffic_key"], job.sm_count,
                                         clocks.get("sm_mhz")),
            "e2e": {"value": w["e2e_value"], "unit": "GCUPS", "h2d_bytes_per_step": w["pairs"].h2d_bytes * world,
                    "d2h_bytes_per_step": (n * 40 + w["n_ops_total"] * 8) * world, "steps": args.steps},
            "gpu_launches": int(w["launches"]),
            "clocks": clocks,
        }
        if int32_value is not None:
            line["config"]["int32_kernel_value"] = int32_value
            line["config"]["dtype_note"] = ("dtype int32 = the arithmetic the results are defined in; lattices inside "
                                            "the host-checked 16-bit bound run two per register (route), all others "
                                            "on the int32 kernels at int32_kernel_value GCUPS on this same batch")
        if extra:
            line["workloads"] = extra
        if strong_block:
            line["strong_scaling"] = strong_block
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = check_against_reference(args.model, w, max(1, args.cpu_pairs or W["cpu_pairs"]))
        if world == 1 and not args.no_cli and args.model == "affine:local":
            eng.close()   # the CLI is its own process: give it the GPU (pool memory, context time slices)
            torch.cuda.empty_cache()
            line["cli"] = cli_leg()
        print(json.dumps(line))
    eng.close()
    if world > 1:
        job.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="affine:local", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0,
                    help="pairs per GPU per step (default: 10k affine:local, 1k est2genome -- BASELINE configs)")
    ap.add_argument("--qlen", type=int, default=0, help="query length (default: the workload's, 1000)")
    ap.add_argument("--tlen", type=int, default=0, help="target length (default: the workload's, 100000)")
    ap.add_argument("--cpu-pairs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the exonerate_b200 CLI leg (N=1 only)")
    ap.add_argument("--only-main", action="store_true", help="skip the other workloads / strong-scaling blocks")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --pairs (default: the workload's batch) is the JOB's total, split over the ranks")
    args = ap.parse_args()
    args.qlen = args.qlen or WORKLOADS[args.model].get("qlen", 1000)
    args.tlen = args.tlen or WORKLOADS[args.model].get("tlen", 100000)
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
