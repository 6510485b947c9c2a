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

from dataclasses import dataclass
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


class RecordView:
    """``DomainBatch.records`` of one or more batches as a read-only sequence that is only built
    when somebody looks at it: result objects carry the row identity of 10^5 domains without
    paying for 10^5 DomainRecord objects on every call (150 ms at 60 601 domains)."""

    def __init__(self, *batches):
        self._batches = batches
        self._list = None

    def _get(self):
        if self._list is None:
            out = []
            for b in self._batches:
                out.extend(b.records)
            self._list = out
        return self._list

    def __len__(self):
        return sum(len(b) for b in self._batches)

    def __iter__(self):
        return iter(self._get())

    def __getitem__(self, k):
        return self._get()[k]

    def __eq__(self, other):
        return list(self) == list(other)

    def __repr__(self):
        return f"RecordView({len(self)} domains)"


class DomainBatch:
    """The flattened collection the engine packs.  Two interchangeable forms:

    * list form  — ``sequences`` (list of str) and ``records`` given up front (hand-built batches,
      FASTA input);
    * array form — what ``extract`` produces: every protein's sequence ONCE in ``blob`` (uint8) and
      per domain a slice (``starts``, ``lengths``) into it, i.e. ``BGCDomain.get_sequence()``
      (BGClib/BGClib.py:3240-3241) as numbers instead of 10^5 Python string slices; K1 cuts the
      slices on the device (``bgca_pack_spans``).  ``sequences`` and ``records`` are then built
      only when somebody asks for them.

    Either way: ``bgc_ids`` / ``bgc_of_seq`` (row identity of the BGC matrix), ``type_names`` /
    ``type_of_seq`` (domain type keys), ``skipped`` (human-readable notes: empty domains ...)."""

    def __init__(self, sequences=None, records=None, bgc_ids=None, bgc_of_seq=None, type_names=None,
                 type_of_seq=None, skipped=None, *, blob=None, starts=None, lengths=None, meta=None):
        self._sequences = None if sequences is None else list(sequences)
        self._records = None if records is None else list(records)
        self.bgc_ids = list(bgc_ids if bgc_ids is not None else [])
        self.bgc_of_seq = np.asarray(bgc_of_seq if bgc_of_seq is not None else [], dtype=np.int32)
        self.type_names = list(type_names if type_names is not None else [])
        self.type_of_seq = np.asarray(type_of_seq if type_of_seq is not None else [], dtype=np.int32)
        self.skipped = list(skipped) if skipped is not None else []
        self._blob = None if blob is None else np.ascontiguousarray(blob, dtype=np.uint8)
        self._starts = None if starts is None else np.ascontiguousarray(starts, dtype=np.int64)
        self._lengths = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32)
        # array form: (protein identifiers, domain ids, ali_from, ali_to) per domain, for lazy records
        self._meta = meta
        if self._sequences is None and self._blob is None:
            raise ValueError("DomainBatch needs sequences or (blob, starts, lengths)")

    def __len__(self):
        return len(self._sequences) if self._sequences is not None else len(self._lengths)

    # ---- the two views of the residues ----------------------------------------------------------
    @property
    def sequences(self) -> List[str]:
        if self._sequences is None:
            text = self._blob.tobytes().decode("ascii")
            self._sequences = [text[a:a + l] for a, l in zip(self._starts.tolist(), self._lengths.tolist())]
        return self._sequences

    def spans(self):
        """(blob uint8, starts int64[n], lengths int32[n]): sequence i = blob[starts[i] : starts[i] + lengths[i]]."""
        if self._blob is None:
            lens = np.fromiter((len(x) for x in self._sequences), dtype=np.int64, count=len(self._sequences))
            starts = np.zeros(len(lens), dtype=np.int64)
            np.cumsum(lens[:-1], out=starts[1:])
            try:
                raw = "".join(self._sequences).encode("ascii")
            except UnicodeEncodeError:
                raise ValueError("sequences must be ASCII protein letters") from None
            self._blob = np.frombuffer(raw, dtype=np.uint8) if raw else np.zeros(0, np.uint8)
            self._starts, self._lengths = starts, lens.astype(np.int32)
        return self._blob, self._starts, self._lengths

    def record_view(self) -> "RecordView":
        """``records`` without building them until they are looked at."""
        return RecordView(self)

    @property
    def records(self) -> List["DomainRecord"]:
        if self._records is None:
            prot, dom, a0, a1 = self._meta
            tn, bi = self.type_names, self.bgc_ids
            self._records = [DomainRecord(i, bi[b], prot[i], dom[i], tn[t], int(a0[i]), int(a1[i]), int(l))
                             for i, (b, t, l) in enumerate(zip(self.bgc_of_seq.tolist(), self.type_of_seq.tolist(),
                                                               self._lengths.tolist()))]
        return self._records

    def digest_update(self, h) -> None:
        """Feeds the residues (in order, with their boundaries) to a hashlib object."""
        blob, starts, lengths = self.spans()
        h.update(np.ascontiguousarray(lengths).tobytes())
        if len(lengths) and (starts[1:] == starts[:-1] + lengths[:-1]).all() and starts[0] == 0 and \
                int(starts[-1] + lengths[-1]) == len(blob):
            h.update(blob.tobytes())                     # contiguous: no copy of the slices needed
        else:
            h.update("\x00".join(self.sequences).encode("ascii"))

    def select_bgcs(self, ids: Sequence[str]) -> "DomainBatch":
        """The batch restricted to the BGCs named in ``ids``, in that order (the reading a
        ``--bgclist`` file gets, BGCtoolkit.py:331-358); unknown ids raise KeyError."""
        pos = {b: k for k, b in enumerate(self.bgc_ids)}
        new_of_old = np.full(len(self.bgc_ids), -1, dtype=np.int32)
        for k, b in enumerate(ids):
            new_of_old[pos[b]] = k
        keep = np.nonzero(new_of_old[self.bgc_of_seq] >= 0)[0] if len(self) else np.zeros(0, np.int64)
        old = self.records
        recs = [DomainRecord(k, *(getattr(old[i], f) for f in
                                 ("bgc_id", "protein_identifier", "domain_id", "type_key", "ali_from", "ali_to", "length")))
                for k, i in enumerate(keep.tolist())]
        seqs = self.sequences
        return DomainBatch([seqs[i] for i in keep.tolist()], recs, list(ids),
                           new_of_old[self.bgc_of_seq[keep]].astype(np.int32), list(self.type_names),
                           self.type_of_seq[keep].astype(np.int32), list(self.skipped))

    # ---- on-disk form (resumable / repeated runs: a database is extracted once) -------------
    def save(self, path) -> None:
        """One .npz: ASCII residues + offsets + type/BGC ids + the identity tables."""
        seqs = self.sequences
        offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in seqs], out=offsets[1:])
        recs = self.records
        np.savez_compressed(
            path, ascii=np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8),
            offsets=offsets, type_of_seq=self.type_of_seq, bgc_of_seq=self.bgc_of_seq,
            type_names=np.asarray(self.type_names, dtype=object), bgc_ids=np.asarray(self.bgc_ids, dtype=object),
            protein=np.asarray([r.protein_identifier for r in recs], dtype=object),
            domain=np.asarray([r.domain_id for r in recs], dtype=object),
            ali=np.asarray([[r.ali_from, r.ali_to] for r in recs], dtype=np.int64).reshape(-1, 2))

    @classmethod
    def load(cls, path) -> "DomainBatch":
        z = np.load(path, allow_pickle=True)
        off = z["offsets"]
        type_names, bgc_ids = list(z["type_names"]), list(z["bgc_ids"])
        ali = z["ali"].reshape(-1, 2)
        meta = ([str(x) for x in z["protein"]], [str(x) for x in z["domain"]], ali[:, 0], ali[:, 1])
        return cls(None, None, bgc_ids, z["bgc_of_seq"].astype(np.int32), type_names,
                   z["type_of_seq"].astype(np.int32), blob=z["ascii"], starts=off[:-1],
                   lengths=np.diff(off).astype(np.int32), meta=meta)


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
    """Flatten ``source`` into a DomainBatch (array form: every protein's sequence once, a domain
    = a slice of it; see DomainBatch).

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

    One pass over the objects that only collects numbers and references: per protein its sequence
    (once), per domain (protein index, ali_from, ali_to, type id, BGC id).  The slice
    ``protein.sequence[ali_from:ali_to]`` (BGClib/BGClib.py:3240-3241) is then evaluated for all
    domains at once with Python's slice rules (negative and out-of-range coordinates included) and
    cut on the device by K1; no per-domain string or record object is made here.
    """
    if unit not in ("domain", "protein"):
        raise ValueError("unit must be 'domain' or 'protein'")
    keep = set(domain_types) if domain_types is not None else None

    parts: List[str] = []          # residue buffer, piecewise: protein sequences / stand-alone strings
    part_len: List[int] = []
    part_of = {}                   # id(protein) -> index in parts
    d_part: List[int] = []         # per domain: which part, raw slice coordinates, ids
    d_a0: List[int] = []
    d_a1: List[int] = []
    d_type: List[str] = []         # type key / BGC name; ids are handed out after empty slices are dropped
    d_bgc: List[str] = []
    d_prot: List[str] = []
    d_dom: List[str] = []
    bgc_ids: List[str] = []
    bgc_index = {}
    type_names: List[str] = []
    type_index = {}

    def bgc_slot(name: str) -> int:
        k = bgc_index.get(name)
        if k is None:
            k = bgc_index[name] = len(bgc_ids)
            bgc_ids.append(name)
        return k

    def type_slot(name: str) -> int:
        k = type_index.get(name)
        if k is None:
            k = type_index[name] = len(type_names)
            type_names.append(name)
        return k

    def part_slot(key, seq: str) -> int:
        k = part_of.get(key)
        if k is None:
            k = part_of[key] = len(parts)
            parts.append(seq)
            part_len.append(len(seq))
        return k

    def add(part: int, a0: int, a1: int, bgc: str, prot_id: str, dom_id: str, tkey: str):
        if keep is not None and tkey not in keep:
            return
        d_part.append(part); d_a0.append(a0); d_a1.append(a1)
        d_type.append(tkey); d_bgc.append(bgc)
        d_prot.append(prot_id); d_dom.append(dom_id)

    def add_string(s_: str, name: str):
        if keep is not None and "sequence" not in keep:
            return
        seq = _norm(s_)
        add(part_slot(("s", len(parts)), seq), 0, len(seq), name, name, "", "sequence")

    def add_protein(p, bgc_name: Optional[str]):
        name = bgc_name if bgc_name is not None else (getattr(p, "parent_cluster_id", "") or p.identifier)
        seq = p.sequence
        pid = p.identifier
        if unit == "protein":
            tkey = getattr(p, "protein_type", "") or "protein"
            if keep is None or tkey in keep:
                add(part_slot(id(p), seq), 0, len(seq), name, pid, "", tkey)
            return
        part = -1
        for d in (merged_domain_view(p) if merge_split_domains else p.domain_list):
            # type key: external alias, then the domain's own alias, then its ID (type_key())
            tkey = alias.get(d.ID) if alias else None
            if tkey is None:
                tkey = getattr(d, "alias", "") or d.ID
            if keep is not None and tkey not in keep:
                continue
            if part < 0:
                part = part_slot(id(p), seq)
            add(part, d.ali_from, d.ali_to, name, pid, d.ID, tkey)

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
        for name, s_ in source.items():
            add_string(s_, str(name))
    else:
        for i, item in enumerate(source):
            if _is_bgc(item):
                add_bgc(item)
            elif _is_protein(item):
                add_protein(item, None)
            elif _is_domain(item):
                p = item.protein
                name = getattr(p, "parent_cluster_id", "") or p.identifier
                tkey = type_key(item, alias)
                if keep is None or tkey in keep:
                    add(part_slot(id(p), p.sequence), item.ali_from, item.ali_to, name, p.identifier, item.ID, tkey)
            elif isinstance(item, str):
                add_string(item, f"seq{i}")
            else:
                raise TypeError(f"cannot extract sequences from {type(item).__name__}")

    # ---- all slices at once, with Python's slice rules ------------------------------------------
    n_all = len(d_part)
    plen = np.asarray(part_len, dtype=np.int64)
    pstart = np.zeros(len(parts), dtype=np.int64)
    if len(parts) > 1:
        np.cumsum(plen[:-1], out=pstart[1:])
    part = np.asarray(d_part, dtype=np.int64)
    L = plen[part] if n_all else np.zeros(0, np.int64)
    a0 = np.asarray(d_a0, dtype=np.int64)
    a1 = np.asarray(d_a1, dtype=np.int64)
    lo = np.clip(np.where(a0 < 0, a0 + L, a0), 0, L)
    hi = np.clip(np.where(a1 < 0, a1 + L, a1), 0, L)
    length = np.maximum(hi - lo, 0)
    ok = length > 0
    skipped = [f"{d_prot[i]}:{d_dom[i]}[{d_a0[i]}:{d_a1[i]}] has an empty sequence" for i in np.nonzero(~ok)[0].tolist()]
    sel = np.nonzero(ok)[0]
    try:
        raw = "".join(parts).encode("ascii")
    except UnicodeEncodeError:
        raise ValueError("sequences must be ASCII protein letters") from None
    blob = np.frombuffer(raw, dtype=np.uint8) if raw else np.zeros(0, np.uint8)
    if len(sel) != n_all:
        sl = sel.tolist()
        d_prot, d_dom = [d_prot[i] for i in sl], [d_dom[i] for i in sl]
        d_type, d_bgc = [d_type[i] for i in sl], [d_bgc[i] for i in sl]
    prot, dom = d_prot, d_dom
    # ids in order of first appearance among the domains that stay (BGCs of a collection got
    # their rows above, domains or not)
    type_ids = np.fromiter((type_slot(t) for t in d_type), dtype=np.int32, count=len(d_type))
    bgc_of = np.fromiter((bgc_slot(b) for b in d_bgc), dtype=np.int32, count=len(d_bgc))
    return DomainBatch(None, None, bgc_ids, bgc_of, type_names, type_ids, skipped, blob=blob,
                       starts=(pstart[part] + lo)[sel], lengths=length[sel].astype(np.int32),
                       meta=(prot, dom, a0[sel], a1[sel]))


def _norm(s: str) -> str:
    """Same normalisation as the reference's sequence setter (BGClib/BGClib.py:1906)."""
    return s.replace("\n", "").strip().replace(" ", "").upper()
