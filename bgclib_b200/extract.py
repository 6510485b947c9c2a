"""Walks unchanged (duck-typed) BGClib objects and flattens them into the arrays
the engine packs.  Nothing is attached to, or mutated on, the reference objects
(they get pickled by BGCtoolkit.py:1249,1255,1354).

Reference contract followed:
  * unit aligned = ``BGCDomain.get_sequence()`` = ``protein.sequence[ali_from:ali_to]``
    (BGClib/BGClib.py:3240-3241); the sequence setter already upper-cased it
    (BGClib/BGClib.py:1903-1907)
  * container walk: ``BGCCollection.bgcs`` (dict id -> BGC, insertion order,
    BGClib/BGClib.py:755) -> ``BGC.protein_list`` (:899) -> ``BGCProtein.domain_list``
    (:1881, ordered by ali_from after filter_domains :2132)
  * ``cbp_only``: only proteins listed in ``BGC.CBPcontent`` (type -> [BGCProtein],
    :889-891), the selection ``save_fasta`` makes (BGCtoolkit.py:1431-1436)
  * domain type key: external alias dict first, then the domain's own alias,
    then its ID — the label precedence of ``domain_string``
    (BGClib/BGClib.py:1991-2002) and the ID-or-alias match of BGCtoolkit.py:1511-1513
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence

import numpy as np


@dataclass(frozen=True)
class DomainRecord:
    """Row/column identity of the domain-pair matrix."""
    index: int
    bgc_id: str
    protein_identifier: str
    domain_id: str
    type_key: str
    ali_from: int
    ali_to: int
    length: int


@dataclass
class DomainBatch:
    sequences: List[str]
    records: List[DomainRecord]
    bgc_ids: List[str]                     # row/column identity of the BGC matrix
    bgc_of_seq: np.ndarray                 # int32[n]
    type_names: List[str]
    type_of_seq: np.ndarray                # int32[n]
    skipped: List[str] = field(default_factory=list)   # human-readable notes (empty domains ...)

    def __len__(self):
        return len(self.sequences)

    def select_bgcs(self, ids: Sequence[str]) -> "DomainBatch":
        """The batch restricted to the BGCs named in ``ids``, in that order (the reading a
        ``--bgclist`` file gets, BGCtoolkit.py:331-358); unknown ids raise KeyError."""
        pos = {b: k for k, b in enumerate(self.bgc_ids)}
        new_of_old = np.full(len(self.bgc_ids), -1, dtype=np.int32)
        for k, b in enumerate(ids):
            new_of_old[pos[b]] = k
        keep = [i for i in range(len(self.sequences)) if new_of_old[self.bgc_of_seq[i]] >= 0]
        recs = [DomainRecord(k, *(getattr(self.records[i], f) for f in
                                 ("bgc_id", "protein_identifier", "domain_id", "type_key", "ali_from", "ali_to", "length")))
                for k, i in enumerate(keep)]
        return DomainBatch([self.sequences[i] for i in keep], recs, list(ids),
                           new_of_old[self.bgc_of_seq[keep]].astype(np.int32), list(self.type_names),
                           self.type_of_seq[keep].astype(np.int32), list(self.skipped))

    # ---- on-disk form (resumable / repeated runs: a database is extracted once) -------------
    def save(self, path) -> None:
        """One .npz: ASCII residues + offsets + type/BGC ids + the identity tables."""
        offsets = np.zeros(len(self.sequences) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in self.sequences], out=offsets[1:])
        np.savez_compressed(
            path, ascii=np.frombuffer("".join(self.sequences).encode("ascii"), dtype=np.uint8),
            offsets=offsets, type_of_seq=self.type_of_seq, bgc_of_seq=self.bgc_of_seq,
            type_names=np.asarray(self.type_names, dtype=object), bgc_ids=np.asarray(self.bgc_ids, dtype=object),
            protein=np.asarray([r.protein_identifier for r in self.records], dtype=object),
            domain=np.asarray([r.domain_id for r in self.records], dtype=object),
            ali=np.asarray([[r.ali_from, r.ali_to] for r in self.records], dtype=np.int64).reshape(-1, 2))

    @classmethod
    def load(cls, path) -> "DomainBatch":
        z = np.load(path, allow_pickle=True)
        blob = z["ascii"].tobytes().decode("ascii")
        off = z["offsets"]
        seqs = [blob[off[i]:off[i + 1]] for i in range(len(off) - 1)]
        type_names, bgc_ids = list(z["type_names"]), list(z["bgc_ids"])
        recs = [DomainRecord(i, bgc_ids[z["bgc_of_seq"][i]], str(z["protein"][i]), str(z["domain"][i]),
                             type_names[z["type_of_seq"][i]], int(z["ali"][i][0]), int(z["ali"][i][1]), len(seqs[i]))
                for i in range(len(seqs))]
        return cls(seqs, recs, bgc_ids, z["bgc_of_seq"].astype(np.int32), type_names,
                   z["type_of_seq"].astype(np.int32))


@dataclass(frozen=True)
class DomainSpan:
    """A domain as the aligner sees it: either one BGCDomain or the merge of consecutive
    fragments of a split domain.  Duck-types the BGCDomain attributes ``extract`` reads."""
    protein: object
    ID: str
    alias: str
    ali_from: int
    ali_to: int
    hmm_from: int
    hmm_to: int
    score: float
    n_fragments: int = 1

    def get_sequence(self) -> str:
        return self.protein.sequence[self.ali_from:self.ali_to]


def split_domain_groups(domain_list: Sequence) -> List[List[int]]:
    """Positions (in ``domain_list``) of the fragment runs that the reference's
    ``fix_core_split_domains`` (BGCtoolkit.py:575-687) stitches into one domain: consecutive
    domains with the same ID whose HMM coordinates advance (prev.hmm_to < next.hmm_from).
    Restated with the reference's exact scan, including its edge behaviour: a run is only opened
    by the FIRST domain of an ID stretch (look-ahead at :601-604), a fragment that does not
    advance is skipped while the run stays open (:609-611), and a run still open when the last
    domain has the same ID but does not advance is dropped (:618-621)."""
    groups: List[List[int]] = []
    n = len(domain_list)
    if n == 0:
        return groups
    run: List[int] = []
    current_id = ""                      # the reference starts from the empty string (:591)
    for k in range(n - 1):
        d = domain_list[k]
        if current_id != d.ID:
            if run:
                groups.append(run)
            run = []
            nxt = domain_list[k + 1]
            if d.ID == nxt.ID and d.hmm_to < nxt.hmm_from:
                run = [k]
            current_id = d.ID
        elif run:
            if domain_list[run[-1]].hmm_to < d.hmm_from:
                run.append(k)
    last = domain_list[n - 1]
    if current_id != last.ID:
        if run:
            groups.append(run)
    elif run:
        if domain_list[run[-1]].hmm_to < last.hmm_from:
            run.append(n - 1)
            groups.append(run)
    return groups


def merged_domain_view(protein) -> List[DomainSpan]:
    """``protein.domain_list`` as it would read after ``fix_core_split_domains``
    (BGCtoolkit.py:662-687) — computed on the side, the protein is NOT modified: per run the
    fragments are removed, the merged domain (first.ali_from .. last.ali_to, first.hmm_from ..
    last.hmm_to, score = sum of the parts) is appended and the list is re-sorted by ali_from
    (stable), run after run, as the reference does."""
    doms = list(protein.domain_list)
    view = [DomainSpan(protein, d.ID, getattr(d, "alias", ""), d.ali_from, d.ali_to, getattr(d, "hmm_from", 0),
                       getattr(d, "hmm_to", 0), getattr(d, "score", 0)) for d in doms]
    spans = list(view)
    for run in split_domain_groups(doms):
        parts = [view[k] for k in run]
        first, last = parts[0], parts[-1]
        merged = DomainSpan(protein, first.ID, first.alias, first.ali_from, last.ali_to, first.hmm_from,
                            last.hmm_to, sum(x.score for x in parts), len(parts))
        for x in parts:
            for pos, y in enumerate(spans):     # list.remove by identity, like the reference's objects
                if y is x:
                    del spans[pos]
                    break
        spans.append(merged)
        spans.sort(key=lambda x: x.ali_from)
    return spans


def type_key(domain, alias: Optional[dict]) -> str:
    if alias:
        try:
            return alias[domain.ID]
        except KeyError:
            pass
    a = getattr(domain, "alias", "")
    return a if a != "" else domain.ID


def _is_collection(obj):
    return hasattr(obj, "bgcs") and isinstance(getattr(obj, "bgcs"), dict)


def _is_bgc(obj):
    return hasattr(obj, "protein_list") and hasattr(obj, "CBPcontent")


def _is_protein_collection(obj):
    return hasattr(obj, "proteins") and isinstance(getattr(obj, "proteins"), dict) and not _is_bgc(obj)


def _is_protein(obj):
    return hasattr(obj, "domain_list") and hasattr(obj, "sequence")


def _is_domain(obj):
    return hasattr(obj, "ali_from") and hasattr(obj, "get_sequence")


def _bgc_proteins(bgc, cbp_only: bool):
    if not cbp_only:
        return list(bgc.protein_list)
    chosen = {id(p) for plist in bgc.CBPcontent.values() for p in plist}
    # keep the 5'-3' order of protein_list, as save_fasta's callers see it
    return [p for p in bgc.protein_list if id(p) in chosen]


def extract(source, *, unit: str = "domain", cbp_only: bool = True, alias: Optional[dict] = None,
            domain_types: Optional[Iterable[str]] = None, merge_split_domains: bool = False) -> DomainBatch:
    """Flatten ``source`` into a DomainBatch.

    source: BGCCollection | BGC | ProteinCollection | iterable of BGC / BGCProtein /
            BGCDomain / str | dict name -> str.
    unit:   "domain"  — one entry per BGCDomain (get_sequence());
            "protein" — one entry per BGCProtein (whole sequence; the form the
                        reference's own domain FASTA exports load back as,
                        e.g. examples/all_KS_domains.fasta).
    domain_types: optional set of type keys to keep (e.g. {"KS"}).
    merge_split_domains: align the domains a protein would have after the reference's
            ``fix_core_split_domains`` post-processing (BGCtoolkit.py:575-687) without running it
            on (and so without modifying) the objects — see ``merged_domain_view``.
    """
    if unit not in ("domain", "protein"):
        raise ValueError("unit must be 'domain' or 'protein'")
    keep = set(domain_types) if domain_types is not None else None

    seqs: List[str] = []
    recs: List[DomainRecord] = []
    bgc_ids: List[str] = []
    bgc_index = {}
    bgc_of: List[int] = []
    type_names: List[str] = []
    type_index = {}
    type_of: List[int] = []
    skipped: List[str] = []

    def bgc_slot(name: str) -> int:
        if name not in bgc_index:
            bgc_index[name] = len(bgc_ids)
            bgc_ids.append(name)
        return bgc_index[name]

    def type_slot(name: str) -> int:
        if name not in type_index:
            type_index[name] = len(type_names)
            type_names.append(name)
        return type_index[name]

    def add(seq: str, bgc_name: str, prot_id: str, dom_id: str, tkey: str, a0: int, a1: int):
        if keep is not None and tkey not in keep:
            return
        if len(seq) == 0:
            skipped.append(f"{prot_id}:{dom_id}[{a0}:{a1}] has an empty sequence")
            return
        recs.append(DomainRecord(len(seqs), bgc_name, prot_id, dom_id, tkey, a0, a1, len(seq)))
        seqs.append(seq)
        bgc_of.append(bgc_slot(bgc_name))
        type_of.append(type_slot(tkey))

    def add_protein(p, bgc_name: Optional[str]):
        name = bgc_name if bgc_name is not None else (getattr(p, "parent_cluster_id", "") or p.identifier)
        if unit == "protein":
            tkey = getattr(p, "protein_type", "") or "protein"
            add(p.sequence, name, p.identifier, "", tkey, 0, len(p.sequence))
            return
        for d in (merged_domain_view(p) if merge_split_domains else p.domain_list):
            add(d.get_sequence(), name, p.identifier, d.ID, type_key(d, alias), d.ali_from, d.ali_to)

    def add_bgc(bgc):
        bgc_slot(bgc.identifier)   # BGCs without eligible domains still get a (zero) row
        for p in _bgc_proteins(bgc, cbp_only):
            add_protein(p, bgc.identifier)

    if _is_collection(source):
        for bgc in source.bgcs.values():
            add_bgc(bgc)
    elif _is_bgc(source):
        add_bgc(source)
    elif _is_protein_collection(source):
        for p in source.proteins.values():
            add_protein(p, None)
    elif isinstance(source, dict):
        for name, s in source.items():
            add(_norm(s), str(name), str(name), "", "sequence", 0, len(s))
    else:
        for i, item in enumerate(source):
            if _is_bgc(item):
                add_bgc(item)
            elif _is_protein(item):
                add_protein(item, None)
            elif _is_domain(item):
                p = item.protein
                name = getattr(p, "parent_cluster_id", "") or p.identifier
                add(item.get_sequence(), name, p.identifier, item.ID, type_key(item, alias),
                    item.ali_from, item.ali_to)
            elif isinstance(item, str):
                add(_norm(item), f"seq{i}", f"seq{i}", "", "sequence", 0, len(item))
            else:
                raise TypeError(f"cannot extract sequences from {type(item).__name__}")

    return DomainBatch(seqs, recs, bgc_ids, np.asarray(bgc_of, dtype=np.int32), type_names,
                       np.asarray(type_of, dtype=np.int32), skipped)


def _norm(s: str) -> str:
    """Same normalisation as the reference's sequence setter (BGClib/BGClib.py:1906)."""
    return s.replace("\n", "").strip().replace(" ", "").upper()
