"""world_size=2 over gloo on CPU: the multi-rank path's host logic — each rank
plans its own share of the tile list (K6), scores it (here with the oracle
standing in for the kernels), and one all_reduce(max) yields the full matrix on
every rank, identical to the single-rank result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from bgclib_b200 import plan_host, synth

INT32_MIN = int(np.iinfo(np.int32).min)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, seqs, lens_sorted, order, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = len(seqs)
        tiles, info = plan_host(lens_sorted, rank=rank, world=world)
        pi, pj = [], []
        for t in tiles:
            for s in range(t["s_begin"], t["s_end"]):
                for q in (t["qa"], t["qb"]):
                    if q >= 0 and s >= q:
                        pi.append(order[q]); pj.append(order[s])
        sc = oracle.pair_list(seqs, pi, pj, nthreads=2)
        m = torch.full((n, n), INT32_MIN, dtype=torch.int32)
        m[pi, pj] = torch.from_numpy(sc)
        m[pj, pi] = torch.from_numpy(sc)
        own_pairs = torch.tensor([len(pi)], dtype=torch.int64)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        dist.all_reduce(own_pairs, op=dist.ReduceOp.SUM)
        if rank == 0:
            ret["matrix"] = m.numpy().copy()
            ret["pairs"] = int(own_pairs.item())
            ret["share"] = info["cells_computed"]
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo_equals_single_rank():
    col = synth.ks_like(40, lo=60, hi=120, mean=90, sd=15)
    seqs = col.sequences()
    lens = col.lengths.astype(np.int32)
    order = np.argsort(lens, kind="stable").astype(np.int32)
    lens_sorted = lens[order]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), seqs, lens_sorted, order, ret), nprocs=2, join=True)
    full = oracle.all_pairs(seqs)
    assert ret["pairs"] == 40 * 41 // 2
    assert (ret["matrix"] == full).all()
    _, info_all = plan_host(lens_sorted)
    assert abs(ret["share"] / info_all["cells_computed"] - 0.5) < 0.05
