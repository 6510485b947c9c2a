"""Public Python surface: the "Distance calculation" feature BGClib announces
for BGCCollection (BGClib/Readme.md:16-18; the commented-out ``--comparative``
switch, BGCtoolkit.py:157-161), computed on B200 through libbgca.so.

    scores = domain_pair_scores(collection)          # all-vs-all Gotoh scores
    result = bgc_similarity(collection)              # BGC x BGC similarity
    attach()                                         # optional: adds both as methods
                                                     # of the reference's BGCCollection

Defaults (this build's choice where the reference defines nothing; SURVEY.md
App. C/D): BLOSUM62, open -12, extend -1 (Biopython's blastp preset; a gap of
length k scores open + (k-1)*extend), mode "local", CBP proteins only.

Multi-GPU: run one process per GPU under torchrun.  When torch.distributed is
initialised, every rank scores its share of the tile list and the partial
results are combined with ONE NCCL all-reduce(max) at the end; every rank
returns the full result.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .engine import ALNSTAT_DTYPE, Engine, MODE_LOCAL, resolve_mode
from .extract import DomainBatch, DomainRecord, RecordView, extract

_ENGINES = {}

INT32_MIN = int(np.iinfo(np.int32).min)


def get_engine(device: Optional[int] = None) -> Engine:
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    device = int(device)
    if device not in _ENGINES:
        _ENGINES[device] = Engine(device)
    return _ENGINES[device]


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


@dataclass
class DomainPairScores:
    scores: np.ndarray              # int32 [n, n], symmetric, diagonal = self scores
                                    # (query-vs-database: [n_query, n_db])
    index: List[DomainRecord]       # row identity (and column identity when database is None); a lazy
                                    # sequence (RecordView): the records are built when first looked at
    computed: Optional[np.ndarray]  # bool, same shape, when same_type_only masked some pairs, else None
    plan: dict                      # tile-plan statistics of this rank
    db_index: Optional[List[DomainRecord]] = None   # column identity of a query-vs-database run
    rows: Optional[tuple] = None    # (r0, r1) when shard_rows kept only this rank's rows scores[r0:r1]


@dataclass
class BGCSimilarity:
    sim: np.ndarray                 # float64 [B, B]           (query-vs-database: [B_query, B_db])
    best: np.ndarray                # int64   [B, B]            best[a, b] = sum_{d in a} rowbest[d, b]
    self_score: np.ndarray          # int64   [B]               (rows)
    bgc_ids: List[str]              # row identity (and column identity when database is None)
    index: List[DomainRecord]
    plan: dict
    db_bgc_ids: Optional[List[str]] = None          # column identity of a query-vs-database run
    db_self_score: Optional[np.ndarray] = None      # int64 [B_db]
    best_db: Optional[np.ndarray] = None            # int64 [B_db, B_query]: the database -> query direction


def row_block(n: int, rank: int, world: int):
    """Rows [r0, r1) of an n-row result that rank ``rank`` of ``world`` keeps under
    ``shard_rows`` (equal blocks of ceil(n / world) rows; the last ranks may get fewer or none)."""
    blk = -(-n // world)
    r0 = min(n, rank * blk)
    return blk, r0, min(n, r0 + blk)


def _check_batch(b: DomainBatch) -> DomainBatch:
    """Type keys and BGC ids are identities: a hand-built batch with repeated names would be
    merged wrongly with a second batch (query vs database)."""
    if len(set(b.type_names)) != len(b.type_names):
        raise ValueError("DomainBatch.type_names must be unique")
    if len(set(b.bgc_ids)) != len(b.bgc_ids):
        raise ValueError("DomainBatch.bgc_ids must be unique")
    return b


def _as_batch(source, unit, cbp_only, alias, domain_types) -> DomainBatch:
    if isinstance(source, DomainBatch):
        return _check_batch(source)
    return extract(source, unit=unit, cbp_only=cbp_only, alias=alias, domain_types=domain_types)


def _merge(q: DomainBatch, d: DomainBatch):
    """Query batch followed by database batch: one span list over the two residue buffers laid end
    to end, a shared type-id space and the BGC index space [query BGCs | database BGCs] (ids are
    NOT merged across the two sides).  Returns ((blob, starts, lengths), types, bgcs)."""
    type_names = list(q.type_names)
    pos = {t: i for i, t in enumerate(type_names)}
    for t in d.type_names:
        if t not in pos:
            pos[t] = len(type_names)
            type_names.append(t)
    d_types = np.asarray([pos[t] for t in d.type_names], dtype=np.int32)
    types = np.concatenate([q.type_of_seq, d_types[d.type_of_seq] if len(d) else d.type_of_seq]).astype(np.int32)
    bgcs = np.concatenate([q.bgc_of_seq, d.bgc_of_seq + len(q.bgc_ids)]).astype(np.int32)
    (qb, qs, ql), (db, ds, dl) = q.spans(), d.spans()
    spans = (np.concatenate([qb, db]), np.concatenate([qs, ds + len(qb)]), np.concatenate([ql, dl]))
    return spans, types, bgcs


def domain_pair_scores(source, *, database=None, mode="local", matrix="BLOSUM62", open_gap_score=-12,
                       extend_gap_score=-1, unit="domain", cbp_only=True, alias=None,
                       domain_types=None, same_type_only=False, device=None,
                       return_device=False, shard_rows=False) -> DomainPairScores:
    """All-vs-all affine-gap alignment scores of the domain sequences in ``source``.

    Raises ValueError for an empty selection, a zero-length sequence or a letter
    outside ``ARNDCQEGHILKMFPSTWYVBZX*`` (as Biopython's PairwiseAligner would).
    Pairs skipped by ``same_type_only`` hold 0 (local) / INT32_MIN (global) and
    are flagged False in ``computed``.

    ``database``: a second collection — only query x database pairs are scored ("a new set of
    BGCs against a precomputed database", BGClib/Readme.md:18) and ``scores`` is
    [n_query, n_database].

    Under a ``torch.distributed`` process group every rank scores its share of the tile list.
    By default one all-reduce(max) leaves the whole matrix on every rank; ``shard_rows=True``
    uses a reduce-scatter(max) instead — half the NVLink traffic — and returns only this rank's
    row block (``result.rows`` = (r0, r1), ``row_block``), so the device-to-host copies of the
    ranks run in parallel.
    """
    batch = _as_batch(source, unit, cbp_only, alias, domain_types)
    n = len(batch)
    if n == 0:
        raise ValueError("no domain sequences selected")
    pdb = database if isinstance(database, PackedDatabase) else None
    eng = pdb.engine if pdb is not None else get_engine(device)
    eng.set_scoring(matrix, open_gap_score, extend_gap_score, mode)
    torch = eng.torch
    dist, rank, world = _dist()
    db = None
    if pdb is not None:
        # persistent database: only the queries are uploaded and encoded; [database | queries]
        db = pdb.batch
        pos = dict(pdb.type_pos)
        for t in batch.type_names:
            pos.setdefault(t, len(pos))
        q_types = np.asarray([pos[t] for t in batch.type_names], dtype=np.int32)[batch.type_of_seq]
        Bd = len(db.bgc_ids)
        eng.pack_append(batch.sequences, type_of_seq=q_types, bgc_of_seq=batch.bgc_of_seq + Bd,
                        n_bgc_total=Bd + max(1, len(batch.bgc_ids)))
        plan = eng.plan(same_type_only=same_type_only, rank=rank, world=world, cross_only=2)
        ncols = len(db)
    elif database is not None:
        db = _as_batch(database, unit, cbp_only, alias, domain_types)
        if len(db) == 0:
            raise ValueError("no domain sequences selected in the database")
        spans, types, _ = _merge(batch, db)
        eng.pack_spans(*spans, type_of_seq=types if same_type_only else None, n_query=n)
        plan = eng.plan(same_type_only=same_type_only, rank=rank, world=world, cross_only=True)
        ncols = len(db)
    else:
        eng.pack_batch(batch, type_of_seq=batch.type_of_seq if same_type_only else None)
        plan = eng.plan(same_type_only=same_type_only, rank=rank, world=world)
        ncols = n
    rows = (0, n) if shard_rows else None      # a single process holds every row
    if dist is not None and shard_rows:
        blk, r0, r1 = row_block(n, rank, world)
        padded = torch.full((blk * world, ncols), INT32_MIN, dtype=torch.int32, device=eng.tdev)
        eng.score(scores=padded[:n])
        mine = torch.empty((blk, ncols), dtype=torch.int32, device=eng.tdev)
        dist.reduce_scatter_tensor(mine, padded, op=dist.ReduceOp.MAX)
        scores, rows = mine[: r1 - r0], (r0, r1)
    else:
        scores = torch.full((n, ncols), INT32_MIN, dtype=torch.int32, device=eng.tdev)
        eng.score(scores=scores)
        if dist is not None:
            dist.all_reduce(scores, op=dist.ReduceOp.MAX)
    computed = None
    if same_type_only:
        computed = scores != INT32_MIN
        if resolve_mode(mode) == MODE_LOCAL:
            scores = torch.where(computed, scores, torch.zeros_like(scores))
    db_index = None if db is None else db.record_view()
    if return_device:
        return DomainPairScores(scores, batch.record_view(), computed, plan, db_index, rows)
    return DomainPairScores(scores.cpu().numpy(), batch.record_view(),
                            None if computed is None else computed.cpu().numpy(), plan, db_index, rows)


def bgc_similarity(source, *, database=None, matrix="BLOSUM62", open_gap_score=-12, extend_gap_score=-1,
                   unit="domain", cbp_only=True, alias=None, domain_types=None, same_type_only=True,
                   bgc_ids=None, device=None, return_device=False, rounds=1, checkpoint=None,
                   _stop_after_rounds=None) -> BGCSimilarity:
    """BGC x BGC similarity from local domain-pair scores:

        rowbest[d,b] = max{ s(d,e) : e in b, type(e) == type(d) }      (0 if none)
                       (``same_type_only=False`` drops the type condition on purpose, SURVEY.md
                       App. D: every domain of b is then a candidate partner of d)
        best[a,b]    = sum_{d in a} rowbest[d,b]
        self[a]      = sum_{d in a} s(d,d)
        sim[a,b]     = (best[a,b] + best[b,a]) / (self[a] + self[b])    (0/0 -> 0, sim[a,a] = 1)

    The domain-pair matrix is never materialised: the DP kernels' epilogue does the
    segmented max (atomicMax into rowbest), a second kernel the segmented sums.
    ``bgc_ids`` fixes the row order (ids absent from ``source`` get zero rows).
    ``database``: a second collection — rows are the BGCs of ``source``, columns the BGCs of
    ``database``; only query x database domain pairs (and self pairs) are aligned.

    Resumable runs (tiles are independent and max is idempotent): ``rounds`` > 1 scores this
    rank's share of the tile list in that many equal parts; with ``checkpoint`` (a path; one file
    per rank) the partial ``rowbest`` / self scores and the list of finished rounds are written
    after every round and a later call on the same input and parameters continues after the last
    finished round.  The reference's own resume idiom is the pickled collection plus the
    ``attempted_domain_prediction`` flag (BGCtoolkit.py:552-560).
    """
    batch = _as_batch(source, unit, cbp_only, alias, domain_types)
    if isinstance(database, PackedDatabase):
        from .alphabet import resolve_matrix
        if (int(open_gap_score), int(extend_gap_score)) != database.scoring[1:] or \
                not np.array_equal(resolve_matrix(matrix), resolve_matrix(database.scoring[0])):
            raise ValueError("scoring parameters differ from the ones the PackedDatabase was built with")
        return _bgc_similarity_persistent(batch, database, same_type_only, return_device)
    if database is not None:
        return _bgc_similarity_cross(batch, _as_batch(database, unit, cbp_only, alias, domain_types),
                                     matrix, open_gap_score, extend_gap_score, same_type_only, device,
                                     return_device)
    ids = list(batch.bgc_ids)
    bgc_of = batch.bgc_of_seq
    if bgc_ids is not None:
        pos = {b: i for i, b in enumerate(bgc_ids)}
        missing = [b for b in ids if b not in pos]
        if missing:
            raise ValueError(f"bgc_ids lacks BGCs present in source: {missing[:3]}")
        remap = np.asarray([pos[b] for b in ids], dtype=np.int32)
        bgc_of = remap[bgc_of] if len(bgc_of) else bgc_of
        ids = list(bgc_ids)
    B = len(ids)
    n = len(batch)
    if B == 0:
        raise ValueError("no BGCs selected")
    eng = get_engine(device)
    torch = eng.torch
    if n == 0:
        z = np.zeros((B, B))
        return BGCSimilarity(z, z.astype(np.int64), np.zeros(B, np.int64), ids, [], {})
    eng.set_scoring(matrix, open_gap_score, extend_gap_score, "local")
    dist, rank, world = _dist()
    eng.pack_batch(batch, type_of_seq=batch.type_of_seq, bgc_of_seq=bgc_of, n_bgc=B)
    rowbest = torch.zeros((n, B), dtype=torch.int32, device=eng.tdev)
    selfsc = torch.zeros((n,), dtype=torch.int32, device=eng.tdev)
    rounds = max(1, int(rounds))
    ckpt = _Checkpoint(checkpoint, rank, batch, bgc_of, B, matrix, open_gap_score, extend_gap_score,
                       same_type_only, rounds, world) if checkpoint is not None else None
    done, plan = set(), None
    if ckpt is not None:
        done, plan = ckpt.load_into(rowbest, selfsc)    # plan statistics of the finished rounds ride along
    for r in range(rounds):
        if r in done:                       # finished before the interruption: nothing to plan or score
            continue
        # round r of this rank = virtual rank (rank * rounds + r) of (world * rounds)
        part = eng.plan(same_type_only=same_type_only, rank=rank * rounds + r, world=world * rounds)
        plan = part if plan is None else {k: (plan[k] + part[k] if k in _PLAN_SUMS else part[k]) for k in part}
        eng.score(rowbest=rowbest, selfsc=selfsc, type_check=False)
        done.add(r)
        if ckpt is not None:
            ckpt.save(rowbest, selfsc, done, plan)
        if _stop_after_rounds is not None and len(done) >= _stop_after_rounds:
            raise KeyboardInterrupt(f"stopped after {len(done)} of {rounds} rounds (test hook)")
    if dist is not None:
        dist.all_reduce(rowbest, op=dist.ReduceOp.MAX)
        dist.all_reduce(selfsc, op=dist.ReduceOp.MAX)
    sim, best, selfsum = eng.reduce_bgc(rowbest, selfsc)
    if return_device:
        return BGCSimilarity(sim, best, selfsum, ids, batch.record_view(), plan)
    return BGCSimilarity(sim.cpu().numpy(), best.cpu().numpy(), selfsum.cpu().numpy(), ids,
                         batch.record_view(), plan)


_PLAN_SUMS = ("n_tiles", "n_pairs", "cells", "cells_computed", "n_fallback32")


class _Checkpoint:
    """Partial state of a ``bgc_similarity`` run on disk: one ``.npz`` per rank, written to a
    temporary name and renamed, tagged with a digest of the input and of every parameter that
    changes the result so that a stale file is never continued."""

    def __init__(self, path, rank, batch, bgc_of, n_bgc, matrix, open_gap, extend_gap, same_type_only, rounds, world):
        import hashlib
        from pathlib import Path
        self.path = Path(f"{path}.rank{rank}.npz")
        h = hashlib.sha256()
        batch.digest_update(h)
        h.update(np.ascontiguousarray(bgc_of, dtype=np.int32).tobytes())
        h.update(np.ascontiguousarray(batch.type_of_seq, dtype=np.int32).tobytes())
        h.update(np.ascontiguousarray(np.asarray(matrix) if not isinstance(matrix, str) else
                                      np.frombuffer(matrix.encode(), dtype=np.uint8)).tobytes())
        h.update(repr((n_bgc, int(open_gap), int(extend_gap), bool(same_type_only), int(rounds), int(world), int(rank))).encode())
        self.tag = h.hexdigest()

    def load_into(self, rowbest, selfsc):
        if not self.path.exists():
            return set(), None
        z = np.load(self.path, allow_pickle=False)
        if str(z["tag"]) != self.tag or z["rowbest"].shape != tuple(rowbest.shape):
            return set(), None                # different input or parameters: start over
        import json
        import torch
        rowbest.copy_(torch.from_numpy(z["rowbest"]))
        selfsc.copy_(torch.from_numpy(z["selfsc"]))
        return set(int(r) for r in z["done"]), json.loads(str(z["plan"]))

    def save(self, rowbest, selfsc, done, plan):
        import json
        tmp = self.path.with_suffix(".tmp.npz")
        np.savez(tmp, tag=np.asarray(self.tag), rowbest=rowbest.cpu().numpy(), selfsc=selfsc.cpu().numpy(),
                 done=np.asarray(sorted(done), dtype=np.int64), plan=np.asarray(json.dumps(plan)))
        tmp.replace(self.path)


def _bgc_similarity_cross(q: DomainBatch, d: DomainBatch, matrix, open_gap_score, extend_gap_score,
                          same_type_only, device, return_device) -> BGCSimilarity:
    Bq, Bd = len(q.bgc_ids), len(d.bgc_ids)
    if Bq == 0 or Bd == 0 or len(q) == 0 or len(d) == 0:
        raise ValueError("query and database must both hold BGCs with domain sequences")
    spans, types, bgcs = _merge(q, d)
    n, B = len(q) + len(d), Bq + Bd
    eng = get_engine(device)
    torch = eng.torch
    eng.set_scoring(matrix, open_gap_score, extend_gap_score, "local")
    dist, rank, world = _dist()
    # the packer sorts by (type, role, length): without types the single block is [queries | database]
    eng.pack_spans(*spans, type_of_seq=types if same_type_only else None, bgc_of_seq=bgcs, n_bgc=B, n_query=len(q))
    plan = eng.plan(same_type_only=same_type_only, rank=rank, world=world, cross_only=True)
    rowbest = torch.zeros((n, B), dtype=torch.int32, device=eng.tdev)
    selfsc = torch.zeros((n,), dtype=torch.int32, device=eng.tdev)
    eng.score(rowbest=rowbest, selfsc=selfsc, type_check=False)
    if dist is not None:
        dist.all_reduce(rowbest, op=dist.ReduceOp.MAX)
        dist.all_reduce(selfsc, op=dist.ReduceOp.MAX)
    sim, best, selfsum = eng.reduce_bgc(rowbest, selfsc)
    sim_x = sim[:Bq, Bq:].contiguous()
    best_qd = best[:Bq, Bq:].contiguous()
    best_dq = best[Bq:, :Bq].contiguous()
    if not return_device:
        sim_x, best_qd, best_dq, selfsum = (x.cpu().numpy() for x in (sim_x, best_qd, best_dq, selfsum))
    return BGCSimilarity(sim_x, best_qd, selfsum[:Bq], list(q.bgc_ids), RecordView(q, d), plan,
                         list(d.bgc_ids), selfsum[Bq:], best_dq)


class PackedDatabase:
    """A collection packed once and kept in HBM ("a precomputed database", BGClib/Readme.md:18):

        db = bgclib_b200.PackedDatabase(mibig_collection, alias=hmmdb.alias)
        hits = bgclib_b200.bgc_similarity(new_collection, database=db, alias=hmmdb.alias)

    It owns its own engine (the packed residues, the type/BGC tables and the self score of every
    database domain stay on the device); every query call only uploads and encodes the query
    sequences behind it (bgca_pack_append) and aligns query x database pairs.  Scoring parameters
    are fixed at construction — the cached self scores depend on them."""

    def __init__(self, source, *, matrix="BLOSUM62", open_gap_score=-12, extend_gap_score=-1, unit="domain",
                 cbp_only=True, alias=None, domain_types=None, device=None):
        import torch
        self.batch = _as_batch(source, unit, cbp_only, alias, domain_types)
        if len(self.batch) == 0 or len(self.batch.bgc_ids) == 0:
            raise ValueError("the database holds no domain sequences")
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.engine = Engine(int(device))
        self.scoring = (matrix, int(open_gap_score), int(extend_gap_score))
        eng = self.engine
        eng.set_scoring(matrix, open_gap_score, extend_gap_score, "local")
        eng.pack_batch(self.batch, type_of_seq=self.batch.type_of_seq, bgc_of_seq=self.batch.bgc_of_seq,
                       n_bgc=len(self.batch.bgc_ids))
        # self scores of the database domains, once (the similarity is normalised with them)
        idx = torch.arange(len(self.batch), dtype=torch.int32, device=eng.tdev)
        self.selfsc = eng.score_pairs32(idx, idx)
        self.type_pos = {t: i for i, t in enumerate(self.batch.type_names)}

    def __len__(self):
        return len(self.batch)


def _bgc_similarity_persistent(q: DomainBatch, db: PackedDatabase, same_type_only, return_device) -> BGCSimilarity:
    d = db.batch
    Bq, Bd, nd, nq = len(q.bgc_ids), len(d.bgc_ids), len(d), len(q)
    if Bq == 0 or nq == 0:
        raise ValueError("the query holds no BGCs with domain sequences")
    eng = db.engine
    torch = eng.torch
    # query type ids in the database's type-id space (types the database lacks get fresh ids)
    pos = dict(db.type_pos)
    for t in q.type_names:
        pos.setdefault(t, len(pos))
    q_types = np.asarray([pos[t] for t in q.type_names], dtype=np.int32)[q.type_of_seq]
    dist, rank, world = _dist()
    eng.set_scoring(*db.scoring, "local")
    # index spaces on this path: [database | queries], BGCs likewise
    eng.pack_append(q.sequences, type_of_seq=q_types, bgc_of_seq=q.bgc_of_seq + Bd, n_bgc_total=Bd + Bq)
    plan = eng.plan(same_type_only=same_type_only, rank=rank, world=world, cross_only=2)
    n, B = nd + nq, Bd + Bq
    rowbest = torch.zeros((n, B), dtype=torch.int32, device=eng.tdev)
    selfsc = torch.zeros((n,), dtype=torch.int32, device=eng.tdev)
    eng.score(rowbest=rowbest, selfsc=selfsc, type_check=False)
    if dist is not None:
        dist.all_reduce(rowbest, op=dist.ReduceOp.MAX)
        dist.all_reduce(selfsc, op=dist.ReduceOp.MAX)
    selfsc[:nd] = db.selfsc
    sim, best, selfsum = eng.reduce_bgc(rowbest, selfsc)
    sim_x = sim[Bd:, :Bd].contiguous()
    best_qd = best[Bd:, :Bd].contiguous()
    best_dq = best[:Bd, Bd:].contiguous()
    if not return_device:
        sim_x, best_qd, best_dq, selfsum = (x.cpu().numpy() for x in (sim_x, best_qd, best_dq, selfsum))
    return BGCSimilarity(sim_x, best_qd, selfsum[Bd:], list(q.bgc_ids), RecordView(q, d), plan,
                         list(d.bgc_ids), selfsum[:Bd], best_dq)


@dataclass
class DomainPairStats:
    stats: np.ndarray               # structured [npairs]: score, identities, columns, q_start, q_end,
                                    # s_start, s_end, gaps (int32 each; bgca_alnstats, include/bgca.h)
    pairs: np.ndarray               # int32 [npairs, 2]: (query, subject) positions in ``index``
    index: List[DomainRecord]

    @property
    def identity(self) -> np.ndarray:
        """identities / alignment columns (0 for an empty local alignment)."""
        c = self.stats["columns"].astype(np.float64)
        return np.divide(self.stats["identities"], c, out=np.zeros_like(c), where=c > 0)


def domain_pair_stats(source, pairs=None, *, mode="local", matrix="BLOSUM62", open_gap_score=-12,
                      extend_gap_score=-1, unit="domain", cbp_only=True, alias=None, domain_types=None,
                      device=None, min_score=None, same_type_only=False) -> DomainPairStats:
    """Alignment-derived statistics WITHOUT traceback (identities, alignment length, start/end
    coordinates, gap columns) for selected domain pairs — what a BiG-SCAPE-style distance on the
    same domains needs (the reference holds no such code).

    ``pairs``: int array [npairs, 2] of (query, subject) positions in the extraction order;
    default = every pair i < j.  The alignment described for a pair is the lexicographic maximum
    of (score, identities, -columns, q_start, s_start, -s_end, -q_end) over all alignments in
    ``mode`` (include/bgca.h, bgca_pair_stats), so ``score`` equals ``domain_pair_scores``' value.
    Coordinates index the domain sequence (0-based, end-exclusive, like ali_from/ali_to,
    BGClib/BGClib.py:3222-3241).

    ``min_score`` (with ``pairs=None``): two stages on the device — the all-vs-all scoring kernels
    first (optionally ``same_type_only``), then statistics only for the pairs i < j whose score
    reaches ``min_score``: the usual shape of an identity-based distance, which is only wanted
    for pairs that are similar at all."""
    batch = _as_batch(source, unit, cbp_only, alias, domain_types)
    n = len(batch)
    if n == 0:
        raise ValueError("no domain sequences selected")
    known = None
    dist, rank, world = _dist()
    if pairs is None and min_score is not None:
        sc = domain_pair_scores(batch, mode=mode, matrix=matrix, open_gap_score=open_gap_score,
                                extend_gap_score=extend_gap_score, same_type_only=same_type_only,
                                device=device, return_device=True)
        t = sc.scores
        if dist is None:
            # one process: the engine still holds the batch the scoring pass packed; pairs are
            # selected on the device and carry their scores as hints (bgca_select_pairs)
            eng = get_engine(device)
            pairs_t, out = eng.pair_stats_above(t.contiguous(), int(min_score), same_type_only)
            stats = np.ascontiguousarray(out.cpu().numpy()).view(ALNSTAT_DTYPE).reshape(-1)
            return DomainPairStats(stats, pairs_t.cpu().numpy(), batch.record_view())
        keep = t >= int(min_score)
        if sc.computed is not None:
            keep &= sc.computed
        sel = keep.triu(diagonal=1).nonzero()                         # row-major: sorted by (i, j)
        known = t[sel[:, 0], sel[:, 1]].to("cpu").numpy()
        pairs = sel.to("cpu").numpy()
    elif pairs is None:
        iu = np.triu_indices(n, k=1)
        pairs = np.stack(iu, axis=1)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    if pairs.size and (pairs.min() < 0 or pairs.max() >= n):
        raise ValueError("pairs: index out of range")
    eng = get_engine(device)
    eng.set_scoring(matrix, open_gap_score, extend_gap_score, mode)
    eng.pack_batch(batch)
    if dist is None:
        out = eng.pair_stats(pairs[:, 0].copy(), pairs[:, 1].copy())
    else:
        # pairs are independent: rank r takes pairs r, r+world, ...; one all-reduce(sum) over the
        # zero-initialised record table puts every record on every rank
        out = eng.torch.zeros((len(pairs), len(ALNSTAT_DTYPE.names)), dtype=eng.torch.int32, device=eng.tdev)
        mine = np.arange(rank, len(pairs), world)
        if len(mine):
            out[eng.torch.as_tensor(mine, device=eng.tdev)] = eng.pair_stats(
                pairs[mine, 0].copy(), pairs[mine, 1].copy(), None if known is None else known[mine].copy())
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    stats = np.ascontiguousarray(out.cpu().numpy()).view(ALNSTAT_DTYPE).reshape(-1)
    return DomainPairStats(stats, pairs, batch.record_view())


def attach(collection_cls=None):
    """Adds ``domain_pair_scores`` / ``similarity_matrix`` as methods of the
    reference's BGCCollection (BGClib/BGClib.py:750) — state is NOT stored on
    the instances, so they stay picklable (BGCtoolkit.py:1249)."""
    if collection_cls is None:
        from BGClib import BGCCollection as collection_cls   # the real reference, if importable
    collection_cls.domain_pair_scores = lambda self, **kw: domain_pair_scores(self, **kw)
    collection_cls.similarity_matrix = lambda self, **kw: bgc_similarity(self, **kw)
    collection_cls.domain_pair_stats = lambda self, pairs=None, **kw: domain_pair_stats(self, pairs, **kw)
    return collection_cls
