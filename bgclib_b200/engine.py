"""ctypes binding of libbgca.so (include/bgca.h) and the per-device Engine.

PyTorch is plumbing here: it owns the caller-visible device buffers and the
stream; every kernel lives in libbgca.so.  There is no CPU fallback — a
missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from .alphabet import NSYM, resolve_matrix

MODE_LOCAL = 0
MODE_GLOBAL = 1
_MODES = {"local": MODE_LOCAL, "global": MODE_GLOBAL}

BGCA_ERR_INVALID = -1
BGCA_ERR_ALPHABET = -2
BGCA_ERR_EMPTY = -3
BGCA_ERR_CUDA = -4
BGCA_ERR_RANGE = -5

_LIB_PATH = Path(__file__).resolve().parent / "libbgca.so"
_lib = None

# every symbol include/bgca.h declares (tests check the library exports them all)
EXPORTED_SYMBOLS = (
    "bgca_version", "bgca_last_error", "bgca_create", "bgca_destroy", "bgca_set_scoring",
    "bgca_pack", "bgca_pack_spans", "bgca_pack_append", "bgca_packed_bytes", "bgca_get_packed", "bgca_plan", "bgca_plan_host",
    "bgca_score", "bgca_score_pairs32", "bgca_pair_stats", "bgca_pair_stats_hinted", "bgca_select_pairs",
    "bgca_pair_stats_selected", "bgca_reduce_bgc", "bgca_score_all_host",
    "bgca_last_launches", "bgca_last_dp_ms", "bgca_last_pack_ms", "bgca_dpx_probe", "bgca_window4g_start",
)


class Tile(ctypes.Structure):
    _fields_ = [("qa", ctypes.c_int32), ("qb", ctypes.c_int32), ("s_begin", ctypes.c_int32),
                ("s_end", ctypes.c_int32), ("variant", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("cells", ctypes.c_int64)]


class PlanInfo(ctypes.Structure):
    _fields_ = [("n_tiles", ctypes.c_int64), ("n_pairs", ctypes.c_int64), ("cells", ctypes.c_int64),
                ("cells_computed", ctypes.c_int64), ("total_pairs", ctypes.c_int64),
                ("total_cells", ctypes.c_int64), ("n_variants", ctypes.c_int32),
                ("n_fallback32", ctypes.c_int32)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def load_library():
    """Load libbgca.so; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build the CUDA library first "
            "(python -m bgclib_b200.build, or __graft_entry__.build()). "
            "bgclib_b200 has no CPU fallback.")
    lib = ctypes.CDLL(str(_LIB_PATH))
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    P = ctypes.POINTER
    lib.bgca_version.restype = ctypes.c_int
    lib.bgca_last_error.restype = ctypes.c_char_p
    lib.bgca_create.argtypes = [ctypes.c_int, P(vp)]
    lib.bgca_destroy.argtypes = [vp]
    lib.bgca_set_scoring.argtypes = [vp, vp, i32, i32, i32]
    lib.bgca_pack.argtypes = [vp, vp, vp, i32, vp, vp, i32, i32, i32, vp]
    lib.bgca_pack_spans.argtypes = [vp, vp, i64, vp, vp, i32, vp, vp, i32, i32, i32, vp]
    lib.bgca_pack_append.argtypes = [vp, vp, vp, i32, vp, vp, i32, i32, vp]
    lib.bgca_packed_bytes.argtypes = [vp]
    lib.bgca_packed_bytes.restype = i64
    lib.bgca_get_packed.argtypes = [vp, vp, vp, vp, vp]
    lib.bgca_plan.argtypes = [vp, i32, i32, i32, i32, P(PlanInfo)]
    lib.bgca_plan_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64,
                                   P(i64), P(PlanInfo)]
    lib.bgca_score.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.bgca_score_pairs32.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.bgca_pair_stats.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.bgca_pair_stats_hinted.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    lib.bgca_select_pairs.argtypes = [vp, vp, i32, i32, P(i64), vp]
    lib.bgca_pair_stats_selected.argtypes = [vp, vp, vp, vp, vp]
    lib.bgca_reduce_bgc.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.bgca_score_all_host.argtypes = [vp, vp, vp, i32, vp]
    lib.bgca_last_launches.argtypes = [vp]
    lib.bgca_last_dp_ms.argtypes = [vp, P(ctypes.c_float)]
    lib.bgca_last_pack_ms.argtypes = [vp, P(ctypes.c_float)]
    lib.bgca_dpx_probe.argtypes = [ctypes.c_int, P(ctypes.c_double), i32]
    lib.bgca_window4g_start.argtypes = [ctypes.c_uint64, ctypes.c_uint64, P(ctypes.c_uint64)]
    for name in EXPORTED_SYMBOLS:
        getattr(lib, name)   # AttributeError here = header and library out of step
    _lib = lib
    return lib


def _check(rc: int):
    if rc == 0:
        return
    msg = load_library().bgca_last_error().decode("utf-8", "replace")
    if rc in (BGCA_ERR_ALPHABET, BGCA_ERR_EMPTY, BGCA_ERR_INVALID, BGCA_ERR_RANGE):
        raise ValueError(msg)
    raise RuntimeError(msg)


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _t_ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def resolve_mode(mode) -> int:
    if isinstance(mode, str):
        try:
            return _MODES[mode.lower()]
        except KeyError:
            raise ValueError(f"mode must be 'local' or 'global', not {mode!r}") from None
    if mode in (MODE_LOCAL, MODE_GLOBAL):
        return int(mode)
    raise ValueError(f"mode must be 'local' or 'global', not {mode!r}")


def concat_ascii(seqs: Sequence[str]):
    """list[str] -> (uint8 ASCII concatenation, int64 offsets[n+1])."""
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offsets[1:])
    try:
        blob = "".join(seqs).encode("ascii")
    except UnicodeEncodeError:
        raise ValueError("sequences must be ASCII protein letters") from None
    ascii_ = np.frombuffer(blob, dtype=np.uint8) if blob else np.zeros(0, np.uint8)
    return ascii_, offsets


def plan_host(len_sorted, type_sorted=None, role_sorted=None, *, same_type_only=False, cross_only=False,
              n_db_first=0, mode=MODE_LOCAL, max_abs_sub=11, open_gap=-12, extend_gap=-1, rank=0, world=1):
    """Host-only K6 planner (no CUDA).  Returns (tiles structured array, info dict).
    Inputs are in SORTED order: by type, then role (0 query / 1 database), then length."""
    lib = load_library()
    lens = np.ascontiguousarray(len_sorted, dtype=np.int32)
    types = None if type_sorted is None else np.ascontiguousarray(type_sorted, dtype=np.int32)
    roles = None if role_sorted is None else np.ascontiguousarray(role_sorted, dtype=np.int32)
    n_tiles = ctypes.c_int64(0)
    info = PlanInfo()
    args = (_np_ptr(lens), _np_ptr(types), _np_ptr(roles), len(lens), int(same_type_only), int(cross_only),
            int(n_db_first), mode, max_abs_sub, open_gap, extend_gap, rank, world)
    _check(lib.bgca_plan_host(*args, None, 0, ctypes.byref(n_tiles), ctypes.byref(info)))
    tiles = (Tile * max(1, n_tiles.value))()
    _check(lib.bgca_plan_host(*args, ctypes.cast(tiles, ctypes.c_void_p), n_tiles.value,
                              ctypes.byref(n_tiles), ctypes.byref(info)))
    arr = np.ctypeslib.as_array(tiles)[: n_tiles.value].copy() if n_tiles.value else \
        np.zeros(0, dtype=np.ctypeslib.as_array(tiles).dtype)
    return arr, info.as_dict()


@dataclass
class ScoreOutputs:
    scores: object = None    # torch int32 [n, n]
    rowbest: object = None   # torch int32 [n, n_bgc]
    selfsc: object = None    # torch int32 [n]


class Engine:
    """One engine per device per process (one process per GPU under torchrun)."""

    def __init__(self, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("bgclib_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        self.device = int(device)
        _check(self._lib.bgca_create(self.device, ctypes.byref(self._h)))
        self.torch = torch
        self.tdev = torch.device("cuda", self.device)
        self.n = 0
        self.n_bgc = 0
        self.n_query = 0
        self.cross = False
        self.mode = MODE_LOCAL
        self.plan_info = None
        self.set_scoring()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.bgca_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.tdev).cuda_stream)

    # ---- scoring parameters ------------------------------------------------
    def set_scoring(self, matrix="BLOSUM62", open_gap_score=-12, extend_gap_score=-1, mode="local"):
        sub = resolve_matrix(matrix)
        for name, v in (("open_gap_score", open_gap_score), ("extend_gap_score", extend_gap_score)):
            if float(v) != int(v):
                raise ValueError(f"{name} must be an integer (scores are computed exactly)")
        self.mode = resolve_mode(mode)
        self.sub = sub
        _check(self._lib.bgca_set_scoring(self._h, _np_ptr(sub), int(open_gap_score),
                                          int(extend_gap_score), self.mode))

    # ---- K1 ------------------------------------------------------------------
    def pack(self, seqs: Sequence[str], type_of_seq=None, bgc_of_seq=None, n_bgc: int = 0, n_query: int = 0):
        ascii_, offsets = concat_ascii(seqs)
        return self.pack_ascii(ascii_, offsets, type_of_seq, bgc_of_seq, n_bgc, n_query)

    def pack_ascii(self, ascii_, offsets, type_of_seq=None, bgc_of_seq=None, n_bgc: int = 0, n_query: int = 0):
        """ascii_: uint8 numpy array (host) or torch uint8 CUDA tensor; offsets int64[n+1] host.
        n_query > 0: the first n_query sequences are queries, the rest a database."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        types = None if type_of_seq is None else np.ascontiguousarray(type_of_seq, dtype=np.int32)
        bgcs = None if bgc_of_seq is None else np.ascontiguousarray(bgc_of_seq, dtype=np.int32)
        if types is not None and len(types) != n:
            raise ValueError("type_of_seq length mismatch")
        if bgcs is not None and len(bgcs) != n:
            raise ValueError("bgc_of_seq length mismatch")
        on_device = not isinstance(ascii_, np.ndarray)
        if on_device:
            if ascii_.dtype != self.torch.uint8 or not ascii_.is_cuda:
                raise ValueError("device ascii must be a CUDA uint8 tensor")
            ptr = _t_ptr(ascii_.contiguous())
        else:
            ascii_ = np.ascontiguousarray(ascii_, dtype=np.uint8)
            ptr = _np_ptr(ascii_) if ascii_.size else ctypes.c_void_p(1)
        with self.torch.cuda.device(self.tdev):
            _check(self._lib.bgca_pack(self._h, ptr, _np_ptr(offsets), n, _np_ptr(types), _np_ptr(bgcs),
                                       int(n_bgc), int(n_query), int(on_device), self._stream()))
        self.n = n
        self.n_query = int(n_query)
        self.cross = False
        self.n_bgc = int(n_bgc) if bgcs is not None else 0
        self.plan_info = None

    def pack_spans(self, blob, starts, lengths, type_of_seq=None, bgc_of_seq=None, n_bgc: int = 0, n_query: int = 0):
        """Sequences given as slices of one residue buffer (``DomainBatch.spans()``): sequence i =
        blob[starts[i] : starts[i] + lengths[i]].  blob: uint8 numpy array (host) or torch uint8 CUDA
        tensor; starts int64[n], lengths int32[n] host.  The slices are cut by K1 on the device."""
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        n = len(starts)
        if len(lengths) != n:
            raise ValueError("starts and lengths must have one entry per sequence")
        types = None if type_of_seq is None else np.ascontiguousarray(type_of_seq, dtype=np.int32)
        bgcs = None if bgc_of_seq is None else np.ascontiguousarray(bgc_of_seq, dtype=np.int32)
        for a in (types, bgcs):
            if a is not None and len(a) != n:
                raise ValueError("type_of_seq / bgc_of_seq must have one entry per sequence")
        on_device = not isinstance(blob, np.ndarray)
        if on_device:
            if blob.dtype != self.torch.uint8 or not blob.is_cuda:
                raise ValueError("device blob must be a CUDA uint8 tensor")
            blob = blob.contiguous()
            ptr, nbytes = _t_ptr(blob), int(blob.numel())
        else:
            blob = np.ascontiguousarray(blob, dtype=np.uint8)
            ptr, nbytes = (_np_ptr(blob) if blob.size else ctypes.c_void_p(1)), int(blob.size)
        with self.torch.cuda.device(self.tdev):
            _check(self._lib.bgca_pack_spans(self._h, ptr, nbytes, _np_ptr(starts), _np_ptr(lengths), n, _np_ptr(types),
                                             _np_ptr(bgcs), int(n_bgc), int(n_query), int(on_device), self._stream()))
        self.n = n
        self.n_query = int(n_query)
        self.cross = False
        self.n_bgc = int(n_bgc) if bgcs is not None else 0
        self.plan_info = None

    def pack_batch(self, batch, type_of_seq=None, bgc_of_seq=None, n_bgc: int = 0, n_query: int = 0):
        """A DomainBatch through its span form: no per-domain Python strings on the way to K1."""
        return self.pack_spans(*batch.spans(), type_of_seq, bgc_of_seq, n_bgc, n_query)

    def pack_append(self, seqs: Sequence[str], type_of_seq=None, bgc_of_seq=None, n_bgc_total: int = 0):
        """Append a query set behind the batch packed earlier (which stays resident as the database);
        a later call replaces the previous query set.  ORIGINAL indices afterwards: [database | queries]."""
        ascii_, offsets = concat_ascii(seqs)
        n_new = len(offsets) - 1
        types = None if type_of_seq is None else np.ascontiguousarray(type_of_seq, dtype=np.int32)
        bgcs = None if bgc_of_seq is None else np.ascontiguousarray(bgc_of_seq, dtype=np.int32)
        for a in (types, bgcs):
            if a is not None and len(a) != n_new:
                raise ValueError("type_of_seq / bgc_of_seq must have one entry per sequence")
        ascii_ = np.ascontiguousarray(ascii_, dtype=np.uint8)
        n_db = self.n - self.n_query if self.n_query else self.n
        with self.torch.cuda.device(self.tdev):
            _check(self._lib.bgca_pack_append(self._h, _np_ptr(ascii_) if ascii_.size else ctypes.c_void_p(1),
                                              _np_ptr(offsets), n_new, _np_ptr(types), _np_ptr(bgcs),
                                              int(n_bgc_total), 0, self._stream()))
        self.n = n_db + n_new
        self.n_query = n_new
        self.cross = False
        if bgcs is not None:
            self.n_bgc = int(n_bgc_total)
        self.plan_info = None

    def get_packed(self):
        nbytes = int(self._lib.bgca_packed_bytes(self._h))
        codes = np.zeros(nbytes, dtype=np.uint8)
        offs = np.zeros(self.n + 1, dtype=np.int64)
        order = np.zeros(self.n, dtype=np.int32)
        _check(self._lib.bgca_get_packed(self._h, _np_ptr(codes), _np_ptr(offs), _np_ptr(order),
                                         self._stream()))
        return codes, offs, order

    # ---- K6 ------------------------------------------------------------------
    def plan(self, same_type_only: bool = False, rank: int = 0, world: int = 1, cross_only=False) -> dict:
        """cross_only: False | True | 2 (2 = cross pairs + query self pairs only, for a persistent
        database whose self scores the caller already holds)."""
        info = PlanInfo()
        _check(self._lib.bgca_plan(self._h, int(same_type_only), int(cross_only), int(rank), int(world),
                                   ctypes.byref(info)))
        self.cross = bool(cross_only)
        self.plan_info = info.as_dict()
        return self.plan_info

    # ---- K2/K3/K4 + fused K5 epilogue ---------------------------------------
    def score(self, scores=None, rowbest=None, selfsc=None, type_check: bool = False):
        """Launch the DP kernels over this rank's tiles into caller-owned CUDA
        int32 tensors (asynchronous on the current stream)."""
        sshape = (self.n_query, self.n - self.n_query) if self.cross else (self.n, self.n)
        for t, shape in ((scores, sshape), (rowbest, (self.n, self.n_bgc)), (selfsc, (self.n,))):
            if t is not None:
                if t.dtype != self.torch.int32 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != shape:
                    raise ValueError(f"output tensor must be contiguous CUDA int32 of shape {shape}")
        with self.torch.cuda.device(self.tdev):
            _check(self._lib.bgca_score(self._h, _t_ptr(scores), _t_ptr(rowbest), _t_ptr(selfsc),
                                        int(type_check), self._stream()))

    def score_pairs32(self, pi, pj):
        t = self.torch
        pi = t.as_tensor(pi, dtype=t.int32, device=self.tdev).contiguous()
        pj = t.as_tensor(pj, dtype=t.int32, device=self.tdev).contiguous()
        out = t.zeros(pi.numel(), dtype=t.int32, device=self.tdev)
        with t.cuda.device(self.tdev):
            _check(self._lib.bgca_score_pairs32(self._h, _t_ptr(pi), _t_ptr(pj), pi.numel(), _t_ptr(out),
                                                self._stream()))
        return out

    def pair_stats(self, pi, pj, known_scores=None):
        """K7: alignment statistics without traceback for pairs (query pi[p], subject pj[p]) of the
        packed batch, ORIGINAL indices.  Returns a CUDA int32 tensor [npairs, 8] whose columns are
        ``ALNSTAT_FIELDS`` (the bgca_alnstats record).  ``known_scores``: the pairs' alignment
        scores when the caller has them (local mode: a faster kernel; the result is exact anyway)."""
        t = self.torch
        pi = t.as_tensor(pi, dtype=t.int32, device=self.tdev).contiguous()
        pj = t.as_tensor(pj, dtype=t.int32, device=self.tdev).contiguous()
        if pi.shape != pj.shape or pi.dim() != 1:
            raise ValueError("pi and pj must be 1-D and of equal length")
        out = t.zeros((pi.numel(), len(ALNSTAT_FIELDS)), dtype=t.int32, device=self.tdev)
        with t.cuda.device(self.tdev):
            if known_scores is None:
                _check(self._lib.bgca_pair_stats(self._h, _t_ptr(pi), _t_ptr(pj), pi.numel(), _t_ptr(out),
                                                 self._stream()))
            else:
                ks = t.as_tensor(known_scores, dtype=t.int32, device=self.tdev).contiguous()
                if ks.shape != pi.shape:
                    raise ValueError("known_scores must have one entry per pair")
                _check(self._lib.bgca_pair_stats_hinted(self._h, _t_ptr(pi), _t_ptr(pj), _t_ptr(ks), pi.numel(),
                                                        _t_ptr(out), self._stream()))
        return out

    def pair_stats_above(self, scores, min_score: int, same_type_only: bool = False):
        """Statistics for every pair i < j of the packed batch whose entry in ``scores`` (the CUDA
        int32 [n, n] matrix ``score()`` filled) reaches ``min_score``, selected on the device.
        Returns (pairs int32 [m, 2], stats int32 [m, 8]) as CUDA tensors, pairs in row-major order."""
        t = self.torch
        if scores.dtype != t.int32 or not scores.is_cuda or not scores.is_contiguous() or \
                tuple(scores.shape) != (self.n, self.n):
            raise ValueError(f"scores must be a contiguous CUDA int32 tensor of shape {(self.n, self.n)}")
        count = ctypes.c_int64(0)
        with t.cuda.device(self.tdev):
            _check(self._lib.bgca_select_pairs(self._h, _t_ptr(scores), int(min_score), int(same_type_only),
                                               ctypes.byref(count), self._stream()))
            m = int(count.value)
            pairs = t.empty((m, 2), dtype=t.int32, device=self.tdev)
            out = t.zeros((m, len(ALNSTAT_FIELDS)), dtype=t.int32, device=self.tdev)
            _check(self._lib.bgca_pair_stats_selected(self._h, _t_ptr(scores), _t_ptr(pairs) if m else None,
                                                      _t_ptr(out) if m else ctypes.c_void_p(8), self._stream()))
        return pairs, out

    def reduce_bgc(self, rowbest, selfsc):
        t = self.torch
        B = self.n_bgc
        best = t.empty((B, B), dtype=t.int64, device=self.tdev)
        selfsum = t.empty((B,), dtype=t.int64, device=self.tdev)
        sim = t.empty((B, B), dtype=t.float64, device=self.tdev)
        with t.cuda.device(self.tdev):
            _check(self._lib.bgca_reduce_bgc(self._h, _t_ptr(rowbest), _t_ptr(selfsc), _t_ptr(best),
                                             _t_ptr(selfsum), _t_ptr(sim), self._stream()))
        return sim, best, selfsum

    def score_all_host(self, ascii_: np.ndarray, offsets: np.ndarray, out: Optional[np.ndarray] = None):
        """The e2e C-ABI call: host ASCII in, host int32 [n,n] out."""
        ascii_ = np.ascontiguousarray(ascii_, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        if out is None:
            out = np.empty((n, n), dtype=np.int32)
        with self.torch.cuda.device(self.tdev):
            _check(self._lib.bgca_score_all_host(self._h, _np_ptr(ascii_), _np_ptr(offsets), n, _np_ptr(out)))
        self.n = n
        self.n_bgc = 0
        self.n_query = 0
        self.cross = False
        return out

    def last_launches(self) -> int:
        return int(self._lib.bgca_last_launches(self._h))

    def last_pack_ms(self) -> float:
        ms = ctypes.c_float(0)
        _check(self._lib.bgca_last_pack_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def last_dp_ms(self) -> float:
        ms = ctypes.c_float(0)
        _check(self._lib.bgca_last_dp_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)


ALNSTAT_FIELDS = ("score", "identities", "columns", "q_start", "q_end", "s_start", "s_end", "gaps")
ALNSTAT_DTYPE = np.dtype([(k, np.int32) for k in ALNSTAT_FIELDS])


def dpx_probe(device: int = 0) -> dict:
    lib = load_library()
    out = (ctypes.c_double * 20)()
    _check(lib.bgca_dpx_probe(device, out, 20))
    names = ["viaddmnmx_s16x2", "vimnmx3_s16x2", "vimnmx_s16x2", "viadd_16x2", "iadd3", "imad", "prmt",
             "viaddmnmx_s32", "cell_mix_s16x2", "sm_mhz",
             "mix_viaddmnmx+viadd16x2", "mix_viaddmnmx+imad", "mix_vimnmx3+viadd16x2", "mix_viadd16x2+imad",
             "lat_viaddmnmx_cyc", "lat_vimnmx3_cyc", "lat_viadd16x2_cyc", "lat_cell_chain_cyc"]
    return {k: float(v) for k, v in zip(names, out)}
