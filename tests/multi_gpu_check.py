#!/usr/bin/env python
"""Multi-rank parity check (not collected by pytest; needs >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29531 tests/multi_gpu_check.py

Every rank runs the public API under an NCCL process group and compares with the oracle:
all-reduce and reduce-scatter forms of the score matrix, fused BGC similarity, the persistent
database, alignment statistics.  Exits non-zero on any mismatch.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    import oracle
    import bgclib_b200
    from bgclib_b200 import synth
    from bgclib_b200.api import row_block

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    col = synth.mibig_scale(40, seed=5)                      # ~1000 mixed domains in 40 BGCs
    seqs = col.sequences()[:400]
    n = len(seqs)
    want = oracle.all_pairs(seqs)

    full = bgclib_b200.domain_pair_scores(seqs)
    assert (full.scores == want).all(), "all-reduce matrix"
    part = bgclib_b200.domain_pair_scores(seqs, shard_rows=True)
    blk, r0, r1 = row_block(n, rank, world)
    assert part.rows == (r0, r1) and (part.scores == want[r0:r1]).all(), "reduce-scatter rows"
    assert world == 1 or full.plan["n_pairs"] < full.plan["total_pairs"], "the tile list was not sharded"

    # fused similarity on a DomainBatch built by hand from the synthetic arrays
    from bgclib_b200.extract import DomainBatch, DomainRecord
    def batch(lo, hi, bgc_shift=0):
        idx = range(lo, hi)
        bg = col.bgc_of_seq[lo:hi]
        ids = sorted(set(bg.tolist()))
        pos = {b: k for k, b in enumerate(ids)}
        recs = [DomainRecord(k, f"bgc{col.bgc_of_seq[i]}", f"p{i}", "d", tnames[col.type_of_seq[i]], 0,
                             len(seqs_all[i]), len(seqs_all[i])) for k, i in enumerate(idx)]
        return DomainBatch([seqs_all[i] for i in idx], recs, [f"bgc{b}" for b in ids],
                           np.asarray([pos[b] for b in bg], np.int32), list(tnames),
                           col.type_of_seq[lo:hi].astype(np.int32))
    seqs_all = col.sequences()
    # type keys are the identity of a type in the API: the generator labels two length classes "KS"
    tnames = [f"{t}#{k}" for k, t in enumerate(col.type_names)]
    b_all = batch(0, 400)
    sim = bgclib_b200.bgc_similarity(b_all)
    o_sim, o_best, o_self, _ = oracle.bgc_similarity_np(want, b_all.bgc_of_seq, b_all.type_of_seq, len(b_all.bgc_ids))
    assert (sim.best == o_best).all() and (sim.sim == o_sim).all(), "fused similarity"

    # persistent database: queries = first 60 domains, database = the rest
    q, d = batch(0, 60), batch(60, 400)
    one = bgclib_b200.bgc_similarity(q, database=d)
    db = bgclib_b200.PackedDatabase(d)
    two = bgclib_b200.bgc_similarity(q, database=db)
    # oracle statement of the cross semantics: same-side pairs do not exist (except self)
    qd = q.sequences + d.sequences
    sc = oracle.all_pairs(qd)
    side = np.arange(len(qd)) >= len(q)
    masked = np.where((side[:, None] != side[None, :]) | np.eye(len(qd), dtype=bool), sc, 0)
    types = np.concatenate([q.type_of_seq, d.type_of_seq])
    Bq, Bd = len(q.bgc_ids), len(d.bgc_ids)
    bgcs = np.concatenate([q.bgc_of_seq, d.bgc_of_seq + Bq])
    w_sim, w_best, w_self, _ = oracle.bgc_similarity_np(masked, bgcs, types, Bq + Bd)
    for name, res in (("one-shot", one), ("persistent", two)):
        ok = [(res.best == w_best[:Bq, Bq:]).all(), (res.best_db == w_best[Bq:, :Bq]).all(),
              (res.self_score == w_self[:Bq]).all(), (res.db_self_score == w_self[Bq:]).all(),
              (res.sim == w_sim[:Bq, Bq:]).all()]
        assert all(ok), f"query-vs-database, {name}: best/best_db/self/db_self/sim ok = {ok} on rank {rank}"

    pairs = np.stack(np.triu_indices(60, k=1), axis=1)
    st = bgclib_b200.domain_pair_stats(seqs[:60], pairs)
    ow = oracle.stats_list(seqs[:60], pairs[:, 0], pairs[:, 1])
    assert all((st.stats[k] == ow[k]).all() for k in ow.dtype.names), "statistics"
    dist.barrier()
    if rank == 0:
        print(f"multi_gpu_check ok: world={world}, {n} domains, every path bit-exact vs the oracle on every rank")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
