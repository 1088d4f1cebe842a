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
        "kernel": "c4b_jit_fill (closed model compiled for sm_100a at run time: region pass + path pass in the "
                  "alignment box, lattice ring in shared memory) + walk",
        "traffic_key": "c4b_jit_fill<protein2genome>",
        "note": "B_alg=104 B/cell (SURVEY 8d: 4 B x 13 states x C=2, reference row layout); the kernel keeps the "
                "lattice ring in shared memory and writes 13 B/cell only inside the alignment box, so the HBM "
                "roofline does not bind it (it is issue/latency bound, profiles/r01d_generic_jit.md). peak ",
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
                return r["score"]
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
            return helpers.oracle_find_path(model, scoring, helpers.PairBuf(q, t, splice=sp),
                                            region_threshold_cells=0, max_ops=len(q) + len(t) + 8)["score"]
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

    def run(self, queries, targets):
        """Wall time for all pairs, dealt round-robin to the processes."""
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
                scores[wi + j * w] = s
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
def ours(args):
    import torch
    import torch.distributed as dist
    import helpers
    from exonerate_b200 import Batch, Engine, Optimal, PairSet, abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the C4 fill has no CPU fallback)")
    torch.cuda.set_device(local)
    # host-side staging threads of the engine: the ranks of one box share its cores
    os.environ.setdefault("C4B_HOST_THREADS", str(max(1, min(16, (os.cpu_count() or 16) // world))))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from exonerate_b200.models import host_model
    params = helpers.load_params()
    scoring = helpers.load_scoring(params)   # the reference's Submat tables (tests/golden/scoring.json)
    W = WORKLOADS[args.model]
    # closed by the host C layer (csrc/host)
    model, _ = host_model(args.model, query_is_protein=W.get("query_is_protein", False))
    n = args.pairs or W["pairs"]
    queries, targets = GENERATORS[args.model](1000 + rank, n, args.qlen, args.tlen)
    splice = None
    if args.model != "affine:local":
        from exonerate_b200.models import splice_arrays
        splice = [splice_arrays(targets[k]) for k in range(n)]   # host C splice predictor (csrc/host/splice.c)
    if args.model == "protein2genome":
        os.environ.setdefault("C4B_GENERIC_JIT", "1")  # the batch is below the auto-specialise size
    pairs = PairSet([queries[k] for k in range(n)], [targets[k] for k in range(n)], splice=splice)
    cells = pairs.cells

    eng = Engine(local)
    stream = torch.cuda.current_stream()
    eng.lib.c4b_engine_set_stream(eng.h, stream.cuda_stream)
    opt = Optimal(eng, model, scoring)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from exonerate_b200.sharding import gather_records
    # weak scaling: rank r owns pairs [r*n, (r+1)*n) of the job's pair list
    shards = [np.arange(r * n, (r + 1) * n) for r in range(world)]

    # ---- device-resident arm: inputs staged in HBM before the timed region ----
    batch = Batch(eng, model, scoring, pairs, want_path=True)
    batch.run()
    dev_results = torch.as_tensor(batch.device_results(), device="cuda")  # c4b_result[n] in HBM, zero-copy

    def step_resident():
        batch.run()
        if world > 1:  # per-pair result records gathered over NCCL/NVLink (north_star)
            gather_records(dev_results, shards, rank, world)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    stop, lines = threading.Event(), []
    sampler = threading.Thread(target=clocks_sampler, args=(stop, lines, local), daemon=True)
    sampler.start()
    launches0 = eng.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fill_ms = []
    e0.record()
    for _ in range(args.steps):
        step_resident()
        if args.fill_timing:
            fill_ms.append(batch.last_fill_ms())
    e1.record()
    barrier()
    launches = eng.kernel_launches() - launches0
    dev_ms = e0.elapsed_time(e1) / args.steps
    if not fill_ms:
        fill_ms.append(batch.last_fill_ms())
    stop.set()
    sampler.join(timeout=3)
    results, ops = batch.fetch(ops_capacity=n * 4096)
    n_ops_total = sum(results[k].n_ops for k in range(n))
    batch.close()

    # ---- end-to-end arm: host buffers in, host results out, every step ----
    out = ((abi.Result * n)(), np.empty(2 * n * 4096, dtype=np.int32))  # host result buffers, reused
    host_scores = torch.zeros((n, 10), dtype=torch.int32, device="cuda")
    for _ in range(min(args.warmup, 2)):
        opt.find_path_raw(pairs, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        got, _ = opt.find_path_raw(pairs, out=out)
        if world > 1:
            host_scores.copy_(torch.from_numpy(np.frombuffer(got, dtype=np.int32).reshape(n, 10)))
            gather_records(host_scores, shards, rank, world)
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    assert all(got[k].score == results[k].score and got[k].n_ops == results[k].n_ops for k in range(n))

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_cells = cells * world
    value = total_cells / (dev_ms * 1e-3) / 1e9
    e2e_value = total_cells / (e2e_ms * 1e-3) / 1e9

    if rank == 0:
        peak, peak_src = measured_peak()
        kfill = float(np.mean(fill_ms))
        achieved = cells * W["b_alg"] / (kfill * 1e-3) / 1e9
        line = {
            "metric": W["metric"], "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": W["workload"] % (queries.shape[1], targets.shape[1]),
                       "pairs_per_gpu": n, "cells_per_step": total_cells, "seed": 1000,
                       "l2": "inputs (%d MB per GPU) exceed the 126 MB L2" % (pairs.h2d_bytes >> 20),
                       "kernel": W["kernel"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(cells, W["traffic_key"]),
                         "note": W["note"] + peak_src, "fill_kernel_ms": kfill},
            "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": pairs.h2d_bytes * world,
                    "d2h_bytes_per_step": (n * 40 + n_ops_total * 8) * world},
            "gpu_launches": int(launches),
            "clocks": summarise_clocks(lines),
        }
        if world == 1 and not args.no_cpu_baseline:
            kind = cpu_arm_available()
            k = max(1, args.cpu_pairs or W["cpu_pairs"])
            pool = CpuPool(kind, 1, args.model)
            wall, cpu_scores = pool.run(queries[:k], targets[:k])
            pool.close()
            assert cpu_scores == [results[i].score for i in range(k)], "CPU baseline disagrees with the GPU scores"
            line["cpu_baseline"] = {"value": k * queries.shape[1] * targets.shape[1] / wall / 1e9, "unit": "GCUPS", "cores": 1,
                                    "kind": kind,
                                    "sample": "first %d pair(s) of the batch, single-threaded, %.1f s" % (k, wall)}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-pairs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fill-timing", action="store_true", help="read the fill-kernel events every step")
    args = ap.parse_args()
    args.qlen = args.qlen or WORKLOADS[args.model].get("qlen", 1000)
    args.tlen = args.tlen or WORKLOADS[args.model].get("tlen", 100000)
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
