"""Deterministic synthetic domain collections for the BASELINE.json configs that
have no real data (configs 3-5).

Family-structured: per domain type, families of ~50 members; an ancestor is
drawn iid from the BLOSUM62 background composition at the type's model length
(HMM LENG values shipped with the reference,
BGClib/data/Domain_models/core_biosynthetic_protein_domains.hmm), members are
the ancestor with per-site substitutions (p ~ U(0.1, 0.6)) and indels
(0.02/site, geometric length, mean 3).  Standard 20 letters only.  NumPy PCG64,
seed = 20260100 + config number.  The generator reports N, pairs and the exact
sum of L_i*L_j so GCUPS denominators are reproducible.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

AA20 = "ARNDCQEGHILKMFPSTWYV"
BACKGROUND = np.array([
    0.0742162, 0.0516145, 0.0446458, 0.0536260, 0.0246875, 0.0342597, 0.0543119, 0.0741469,
    0.0262130, 0.0679174, 0.0989079, 0.0581557, 0.0249902, 0.0474185, 0.0385380, 0.0572290,
    0.0508914, 0.0130300, 0.0322815, 0.0729191])
BACKGROUND = BACKGROUND / BACKGROUND.sum()
_ASCII = np.frombuffer(AA20.encode(), dtype=np.uint8)

# type -> (mean length, sd); LENG of the shipped HMMs; T/ACP ~ 70 (no model shipped)
TYPE_LENGTHS = {"A": (424, 20), "C": (457, 20), "KS": (253, 12), "KS420": (420, 15), "AT": (319, 15),
                "T/ACP": (70, 4)}


@dataclass
class SynthCollection:
    ascii: np.ndarray            # uint8 concatenation
    offsets: np.ndarray          # int64[n+1]
    type_of_seq: np.ndarray      # int32[n]
    type_names: List[str]
    bgc_of_seq: np.ndarray       # int32[n]
    n_bgc: int
    name: str

    def __len__(self):
        return len(self.offsets) - 1

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    def sequences(self) -> List[str]:
        blob = self.ascii.tobytes().decode("ascii")
        return [blob[self.offsets[i]:self.offsets[i + 1]] for i in range(len(self))]

    def subset(self, n: int) -> "SynthCollection":
        n = min(n, len(self))
        return SynthCollection(self.ascii[: self.offsets[n]], self.offsets[: n + 1], self.type_of_seq[:n],
                               self.type_names, self.bgc_of_seq[:n], self.n_bgc, f"{self.name}[:{n}]")

    def pair_stats(self, same_type_only: bool = False) -> dict:
        """Unique pairs i<=j and the exact cell count sum L_i*L_j over them."""
        L = self.lengths.astype(np.int64)
        pairs = cells = 0
        groups = [np.arange(len(L))] if not same_type_only else \
            [np.nonzero(self.type_of_seq == t)[0] for t in range(len(self.type_names))]
        for g in groups:
            l = L[g]
            s, s2 = int(l.sum()), int((l * l).sum())
            pairs += len(l) * (len(l) + 1) // 2
            cells += (s * s + s2) // 2
        return {"n": len(L), "pairs": int(pairs), "cells": int(cells)}


def _mutate(rng, anc: np.ndarray) -> np.ndarray:
    L = len(anc)
    seq = anc.copy()
    p = rng.uniform(0.1, 0.6)
    mask = rng.random(L) < p
    seq[mask] = rng.choice(20, size=int(mask.sum()), p=BACKGROUND)
    n_indel = rng.poisson(0.02 * L)
    for _ in range(n_indel):
        pos = int(rng.integers(0, len(seq) + 1))
        k = int(rng.geometric(1.0 / 3.0))
        if rng.random() < 0.5:
            seq = np.concatenate([seq[:pos], rng.choice(20, size=k, p=BACKGROUND), seq[pos:]])
        elif len(seq) - k > 8:
            seq = np.concatenate([seq[:pos], seq[pos + k:]])
    return seq


def _family_members(rng, n: int, mean: float, sd: float, lo: int, hi: int, family_size: int):
    out = []
    while len(out) < n:
        L = int(np.clip(round(rng.normal(mean, sd)), lo, hi))
        anc = rng.choice(20, size=L, p=BACKGROUND)
        for _ in range(min(family_size, n - len(out))):
            m = _mutate(rng, anc)
            if len(m) > hi:
                m = m[:hi]
            elif len(m) < lo:
                m = np.concatenate([m, rng.choice(20, size=lo - len(m), p=BACKGROUND)])
            out.append(m.astype(np.uint8))
    return out


def _assemble(codes_list, types, type_names, bgcs, n_bgc, name) -> SynthCollection:
    offsets = np.zeros(len(codes_list) + 1, dtype=np.int64)
    np.cumsum([len(c) for c in codes_list], out=offsets[1:])
    ascii_ = _ASCII[np.concatenate(codes_list)]
    return SynthCollection(np.ascontiguousarray(ascii_), offsets, np.asarray(types, np.int32), type_names,
                           np.asarray(bgcs, np.int32), n_bgc, name)


def ks_like(n: int = 10_000, seed: int = 20260103, mean: float = 420, sd: float = 15, lo: int = 360,
            hi: int = 480, family_size: int = 50, name: Optional[str] = None) -> SynthCollection:
    """Config 3: n KS-like domains, L ~ N(420, 15^2) clipped to [360, 480], one type,
    one BGC per domain."""
    rng = np.random.default_rng(seed)
    seqs = _family_members(rng, n, mean, sd, lo, hi, family_size)
    order = rng.permutation(n)           # families are not contiguous in input order
    seqs = [seqs[i] for i in order]
    return _assemble(seqs, np.zeros(n, np.int32), ["KS"], np.arange(n, dtype=np.int32), n,
                     name or f"synthetic {n} KS-like domains (~{int(mean)} aa)")


def mixed_collection(n_bgc: int, seed: int, type_weights: dict, dom_lo: int = 8, dom_hi: int = 40,
                     lo: int = 70, hi: int = 600, family_size: int = 50, name: str = "") -> SynthCollection:
    """Configs 4/5: n_bgc BGCs, each with U[dom_lo, dom_hi] domains whose types are
    drawn from type_weights; sequences per type are family-structured."""
    rng = np.random.default_rng(seed)
    names = list(type_weights)
    w = np.array([type_weights[k] for k in names], dtype=float)
    w /= w.sum()
    counts = rng.integers(dom_lo, dom_hi + 1, size=n_bgc)
    n = int(counts.sum())
    types = rng.choice(len(names), size=n, p=w).astype(np.int32)
    bgcs = np.repeat(np.arange(n_bgc, dtype=np.int32), counts)
    seqs: list = [None] * n
    for t, tname in enumerate(names):
        idx = np.nonzero(types == t)[0]
        mean, sd = TYPE_LENGTHS[tname]
        members = _family_members(rng, len(idx), mean, sd, lo, hi, family_size)
        perm = rng.permutation(len(idx))
        for dst, src in zip(idx, perm):
            seqs[dst] = members[src]
    # display names; type keys must stay unique (they are the identity of a type in the API)
    tn = ["KS+KS_C" if k == "KS420" else k for k in names]
    return _assemble(seqs, types, tn, bgcs, n_bgc, name)


def mibig_scale(n_bgc: int = 2500, seed: int = 20260104) -> SynthCollection:
    """Config 4: ~60k mixed-length domains (A, C, KS, AT, T/ACP; 70-600 aa) in 2.5k BGCs."""
    return mixed_collection(n_bgc, seed, {"T/ACP": 61, "C": 44, "A": 43, "KS": 15, "KS420": 15, "AT": 14},
                            name=f"synthetic MIBiG-scale {n_bgc} BGCs, mixed domains 70-600 aa")


def nrps_pks_100k(n_bgc: int = 4167, seed: int = 20260105) -> SynthCollection:
    """Config 5: ~100k NRPS/PKS domains skewed long (>= 50 % A/C/KS)."""
    return mixed_collection(n_bgc, seed, {"A": 30, "C": 30, "KS420": 15, "AT": 10, "T/ACP": 15},
                            name=f"synthetic {n_bgc} BGCs ~100k NRPS/PKS domains")


# ---- the same collections as OBJECTS with the reference's attribute surface ---------------------
# (BGCCollection.bgcs -> BGC.protein_list / CBPcontent -> BGCProtein.domain_list -> BGCDomain,
# BGClib/BGClib.py:750-906, 1828-1887, 3222-3241) so that the Python API can be driven end to end
# at configs[3] scale without the reference installed.

class SynthDomain:
    __slots__ = ("protein", "ID", "alias", "ali_from", "ali_to")

    def __init__(self, protein, ID, alias, ali_from, ali_to):
        self.protein, self.ID, self.alias, self.ali_from, self.ali_to = protein, ID, alias, ali_from, ali_to

    def get_sequence(self):                      # BGClib/BGClib.py:3240-3241
        return self.protein.sequence[self.ali_from:self.ali_to]


class SynthProtein:
    def __init__(self, identifier, sequence, parent_cluster_id):
        self.identifier, self.parent_cluster_id = identifier, parent_cluster_id
        self.protein_type, self.role = "", "biosynthetic"
        self.domain_list = []
        self._sequence = sequence

    @property
    def sequence(self):
        return self._sequence


class SynthBGC:
    def __init__(self, identifier):
        self.identifier = identifier
        self.protein_list, self.proteins, self.CBPcontent = [], {}, {}


class SynthBGCCollection:
    def __init__(self):
        self.bgcs, self.name = {}, ""


def as_collection(col: SynthCollection, domains_per_protein: int = 4, seed: int = 1) -> SynthBGCCollection:
    """``col`` as a collection of BGC / protein / domain objects: the domains of a BGC are laid
    out, in order, on proteins of ``domains_per_protein`` domains each with short linkers between
    them.  ``extract(as_collection(col))`` yields col's sequences in col's order."""
    rng = np.random.default_rng(seed)
    text = col.ascii.tobytes().decode("ascii")
    off = col.offsets.tolist()
    out = SynthBGCCollection()
    linkers = ["".join(AA20[k] for k in rng.integers(0, 20, size=int(n))) for n in rng.integers(3, 12, size=64)]
    order = np.argsort(col.bgc_of_seq, kind="stable")
    assert (order == np.arange(len(col))).all(), "domains of a BGC must be contiguous"
    bgcs = col.bgc_of_seq.tolist()
    types = col.type_of_seq.tolist()
    i, n, lk = 0, len(col), 0
    while i < n:
        b = bgcs[i]
        bgc = SynthBGC(f"BGC{b:06d}")
        out.bgcs[bgc.identifier] = bgc
        cds = 0
        while i < n and bgcs[i] == b:
            parts, spans, pos = [], [], 0
            for _ in range(domains_per_protein):
                if i >= n or bgcs[i] != b:
                    break
                link = linkers[lk & 63]
                lk += 1
                seq = text[off[i]:off[i + 1]]
                parts.append(link); parts.append(seq)
                pos += len(link)
                spans.append((col.type_names[types[i]], pos, pos + len(seq)))
                pos += len(seq)
                i += 1
            p = SynthProtein(f"{bgc.identifier}~L0+CDS{cds}", "".join(parts), bgc.identifier)
            p.domain_list = [SynthDomain(p, t, t, a, z) for t, a, z in spans]
            bgc.protein_list.append(p)
            bgc.proteins[p.identifier] = p
            bgc.CBPcontent.setdefault("cbp", []).append(p)
            cds += 1
    return out
