"""CPU tier, world_size 2 over gloo: the N>1 path's host logic -- LPT sharding
of a ragged pair list, per-rank processing of its own shard only, all-gather of
the fixed-size result records and restoration of the caller's pair order.  The
per-pair "compute" here is the CPU oracle (test infrastructure), so the check is
that sharded == serial, record for record."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from exonerate_b200 import abi
from exonerate_b200.sharding import gather_records, lpt_shards

SHAPES = [(40, 900), (10, 30), (300, 300), (25, 2000), (120, 80), (64, 640), (7, 7), (200, 1500), (90, 90)]


def make_pairs():
    return [helpers.dna_pair(5000 + k, ql, tl) for k, (ql, tl) in enumerate(SHAPES)]


def records_for(indices, pairs):
    params = helpers.load_params()
    scoring = helpers.load_scoring(params)
    model, _ = helpers.load_model("affine_local_dna", params)
    rec = np.zeros((len(indices), 10), dtype=np.int32)
    for row, k in enumerate(indices):
        q, t = pairs[k]
        r = helpers.oracle_viterbi(model, scoring, helpers.PairBuf(q, t), abi.MODE_FIND_PATH)
        rec[row, :6] = [r["score"], r["region"][0], r["region"][1], r["region"][0] + r["region"][2],
                        r["region"][1] + r["region"][3], len(r["ops"])]
    return rec


def worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = make_pairs()
    shards = lpt_shards([len(q) * len(t) for q, t in pairs], world)
    local = torch.from_numpy(records_for(shards[rank], pairs))
    full = gather_records(local, shards, rank, world)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_lpt_shards_balance_and_cover():
    costs = [len(q) * len(t) for q, t in make_pairs()]
    for world in (1, 2, 3, 8):
        shards = lpt_shards(costs, world)
        allidx = np.sort(np.concatenate(shards))
        assert list(allidx) == list(range(len(costs)))           # every pair exactly once
        loads = [sum(costs[k] for k in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)              # LPT bound
        assert all(list(s) == sorted(s) for s in shards)


def test_two_rank_gather_equals_serial(tmp_path):
    out = str(tmp_path / "gathered.npy")
    mp.spawn(worker, args=(2, free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    pairs = make_pairs()
    want = records_for(list(range(len(pairs))), pairs)
    assert (got == want).all()
