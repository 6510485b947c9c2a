#!/usr/bin/env python
"""Headline benchmark: all-vs-all Gotoh scoring of a synthetic domain collection.

    python bench.py --gpus N --steps K --warmup W            # this build (B200)
    python bench.py --impl reference --gpus N ...            # CPU arm (oracle port, all host cores)

A step = one pass of the hot path over the whole batch: every pair i<=j of the
packed collection scored (local, BLOSUM62, -12/-1), results written to the
n x n int32 matrix, and — for N>1 — combined with one NCCL reduce-scatter(max)
that leaves rank r with its block of rows (the full matrix, sharded by rows).
N=1 workload: BASELINE.json configs[2] (10k KS-like domains, ~420 aa).  N>1 is
weak scaling: round(10000*sqrt(N)) domains from the same generator, so the cells
per GPU stay constant.  metric = GCUPS = 1e9 DP cells/s over the UNIQUE pairs
with true lengths (padding, the redundant mirrored half-pairs and wind-up lanes
are not credited).

After the timed loop the same process runs, untimed for the headline but reported in the same
JSON line:
  "fused"          north_star's end product on ONE fixed collection at every N (strong scaling):
                   BASELINE.json configs[3] — 60 601 mixed domains (70-600 aa) in 2 500 BGCs,
                   same-type pairs — scored with the fused rowbest epilogue, combined with two
                   NCCL all-reduce(max) and reduced to the BGC x BGC similarity; at N = 8 also
                   one step of configs[4] (99 722 domains, every pair).
  "parity_sample"  every rank checks 2 000 random entries of its score-matrix rows and 300 random
                   rowbest entries of the fused result against the oracle, and rank 0 the whole
                   best / self / sim of a 300-domain sub-collection (oracle = checker only).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N1_DOMAINS = 10_000
SEED = 20260103
ALG_LANE_OPS_PER_CELL = 3.0   # local, s16x2: 6 DPX/ALU ops per 2 cells
ALU_LANES_PER_CLK_SM = 64     # alu pipe: 16 lanes/clk/SMSP x 4 (B300_MICROARCH.md "Pipe rates")
ISSUE_LANES_PER_CLK_SM = 128  # issue limit: one warp instruction per clock per scheduler x 4


def workload(n_gpus: int, n_override: int | None):
    from bgclib_b200 import synth
    n = n_override or int(round(N1_DOMAINS * math.sqrt(n_gpus)))
    col = synth.ks_like(n, seed=SEED)
    return col, f"{n} synthetic KS-like domains (L~N(420,15) in [360,480], seed {SEED}), all pairs i<=j"


def bench_config(wl: str) -> dict:
    """The `config` object — identical keys and values in both arms."""
    return {"workload": wl, "mode": "local", "matrix": "BLOSUM62", "open": -12, "extend": -1}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(col, seconds_target: float, nthreads: int = 0):
    """Times the oracle port (scalar int32 Gotoh, OpenMP over pairs) on a bounded
    prefix of the workload.  Returns (gcups, pairs_per_s, cores, sample description, seconds)."""
    import oracle
    if nthreads <= 0:
        # every host core, also under torchrun (which exports OMP_NUM_THREADS=1 to its workers)
        nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = nthreads
    probe = col.subset(24 * max(1, cores // 8 + 1))
    t0 = time.perf_counter()
    oracle.all_pairs(probe.sequences(), nthreads=nthreads)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = probe.pair_stats()["cells"] / dt
    L2 = float(np.mean(col.lengths)) ** 2
    n = int(min(len(col), max(32, math.sqrt(2.0 * rate * seconds_target / L2))))
    sample = col.subset(n)
    st = sample.pair_stats()
    t0 = time.perf_counter()
    oracle.all_pairs(sample.sequences(), nthreads=nthreads)
    dt = time.perf_counter() - t0
    return st["cells"] / dt / 1e9, st["pairs"] / dt, cores, \
        f"first {n} sequences of the workload ({st['pairs']} pairs, {st['cells']:.3e} cells)", dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    col, wl = workload(args.gpus, args.n_seqs)
    vals, secs = [], []
    desc, cores, pps = "", 0, 0.0
    for i in range(args.warmup + args.steps):
        g, p, cores, desc, dt = cpu_reference_run(col, args.cpu_seconds)
        if i >= args.warmup:
            vals.append(g); secs.append(dt); pps = p
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "GCUPS", "value": v, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.mean(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": bench_config(wl),
        "note": "CPU restatement of Biopython PairwiseAligner.score semantics (oracle port); the reference "
                "has no aligner of its own; each step = a bounded sample of the workload",
        "pairs_per_s": pps,
        "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def _all_sum(dist, dev, *vals):
    """Sum of small integers over ranks (identity without a process group)."""
    import torch
    t = torch.tensor(list(vals), dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]


def matrix_parity_sample(col, result, row0, rng, m=2000):
    """m random entries (i, j) of this rank's rows of the score matrix against the oracle."""
    import oracle
    rows, n = result.shape
    li = rng.integers(0, rows, size=m)
    j = rng.integers(0, n, size=m)
    got = result[li, j].cpu().numpy()
    want = oracle.pair_list(col.sequences(), li + row0, j, nthreads=1)
    return m, int((got != want).sum())


def rowbest_parity_sample(col, rowbest, rng, m=300):
    """m random entries rowbest[d, b] of the combined fused result against their definition
    (SURVEY.md App. D): max of the oracle's local scores of d with the same-type domains of b."""
    import oracle
    n, B = rowbest.shape
    d = rng.integers(0, n, size=m)
    b = rng.integers(0, B, size=m)
    order = np.argsort(col.bgc_of_seq, kind="stable")
    start = np.searchsorted(col.bgc_of_seq[order], np.arange(B + 1))
    pi, pj, seg = [], [], []
    for k in range(m):
        mem = order[start[b[k]]:start[b[k] + 1]]
        mem = mem[col.type_of_seq[mem] == col.type_of_seq[d[k]]]
        pi.extend([d[k]] * len(mem)); pj.extend(mem.tolist()); seg.extend([k] * len(mem))
    want = np.zeros(m, dtype=np.int64)
    if pi:
        sc = oracle.pair_list(col.sequences(), np.asarray(pi), np.asarray(pj), nthreads=1)
        np.maximum.at(want, np.asarray(seg), sc)
    got = rowbest[d, b].cpu().numpy()
    return len(pi), m, int((got != want).sum())


def fused_leg(eng, col, dist, rank, world, dev, steps, warm, same_type_only, check_rng=None):
    """north_star's end product on a fixed collection (strong scaling): zero rowbest/selfsc ->
    DP kernels with the fused rowbest epilogue on this rank's tiles -> all-reduce(max) x2 ->
    BGC reduction.  Timed with CUDA events on the launch stream, max over ranks."""
    import torch
    n, B = len(col), col.n_bgc
    eng.pack_ascii(col.ascii, col.offsets, col.type_of_seq, col.bgc_of_seq, B)
    plan = eng.plan(same_type_only=same_type_only, rank=rank, world=world)
    rowbest = torch.zeros((n, B), dtype=torch.int32, device=dev)
    selfsc = torch.zeros((n,), dtype=torch.int32, device=dev)
    out = {}

    def step():
        rowbest.zero_()
        selfsc.zero_()
        eng.score(rowbest=rowbest, selfsc=selfsc, type_check=False)
        if dist is not None:
            dist.all_reduce(rowbest, op=dist.ReduceOp.MAX)
            dist.all_reduce(selfsc, op=dist.ReduceOp.MAX)
        out["r"] = eng.reduce_bgc(rowbest, selfsc)

    launches = 0
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
        launches += eng.last_launches() + 2
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dp = torch.tensor([eng.last_dp_ms()], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(dp, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    sim = out["r"][0]
    res = {"workload": f"{col.name}: {n} domains, {B} BGCs, "
                       f"{'same-type pairs' if same_type_only else 'every pair'} (seed fixed)",
           "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms,
           "gcups": plan["total_cells"] / (ms * 1e-3) / 1e9, "pairs_per_s": plan["total_pairs"] / (ms * 1e-3),
           "pairs": plan["total_pairs"], "cells": plan["total_cells"], "dp_ms_max_rank": float(dp.item()),
           "collective_bytes": int(rowbest.nbytes + selfsc.nbytes) if world > 1 else 0,
           "collective": "2 x NCCL all-reduce(max) over int32 rowbest[n, B] and selfsc[n]" if world > 1 else None,
           "gpu_launches": launches,
           "sim_diag_ones": bool((torch.diagonal(sim) == 1).all().item()),
           "sim_symmetric": bool((sim == sim.T).all().item())}
    chk = None
    if check_rng is not None:
        chk = rowbest_parity_sample(col, rowbest, check_rng)
    del rowbest, selfsc
    return res, chk


def subcollection_parity(eng, col, dist, rank, world, dev, n_sub=300):
    """Fused similarity of the first n_sub domains through every rank; rank 0 compares best, self
    and sim with the oracle's plain-loop statement.  Returns (entries checked, mismatches)."""
    import torch
    import oracle
    sub = col.subset(n_sub)
    B = int(sub.bgc_of_seq.max()) + 1
    eng.pack_ascii(sub.ascii, sub.offsets, sub.type_of_seq, sub.bgc_of_seq, B)
    eng.plan(same_type_only=True, rank=rank, world=world)
    rowbest = torch.zeros((n_sub, B), dtype=torch.int32, device=dev)
    selfsc = torch.zeros((n_sub,), dtype=torch.int32, device=dev)
    eng.score(rowbest=rowbest, selfsc=selfsc, type_check=False)
    if dist is not None:
        dist.all_reduce(rowbest, op=dist.ReduceOp.MAX)
        dist.all_reduce(selfsc, op=dist.ReduceOp.MAX)
    sim, best, selfsum = eng.reduce_bgc(rowbest, selfsc)
    if rank != 0:
        return 0, 0
    nth = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sc = oracle.all_pairs(sub.sequences(), nthreads=nth)
    o_sim, o_best, o_self, _ = oracle.bgc_similarity_np(sc, sub.bgc_of_seq, sub.type_of_seq, B)
    bad = int((best.cpu().numpy() != o_best).sum() + (selfsum.cpu().numpy() != o_self).sum() +
              (sim.cpu().numpy() != o_sim).sum())
    return 2 * B * B + B, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-seqs", type=int, default=None, help="override the workload size (tuning only)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--probe", action="store_true", help="(default at N=1) run the DPX issue-rate microbenchmark")
    ap.add_argument("--no-probe", action="store_true", help="skip the DPX issue-rate microbenchmark")
    ap.add_argument("--no-fused", action="store_true", help="skip the fused BGC-similarity leg (configs[3])")
    ap.add_argument("--fused-steps", type=int, default=3)
    ap.add_argument("--fused-config5", action="store_true",
                    help="also one step of configs[4] (99 722 domains, every pair); default at N = 8")
    ap.add_argument("--no-parity-sample", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    import bgclib_b200
    from bgclib_b200.api import INT32_MIN
    col, wl = workload(world, args.n_seqs)
    n = len(col)
    eng = bgclib_b200.get_engine(local_rank)
    eng.set_scoring("BLOSUM62", -12, -1, "local")

    # inputs resident in HBM before the timed region: packed batch + this rank's tile list
    eng.pack_ascii(col.ascii, col.offsets)
    plan = eng.plan(same_type_only=False, rank=rank, world=world)
    from bgclib_b200.api import row_block
    blk, r0, r1 = row_block(n, rank, world)
    padded = torch.empty((blk * world, n), dtype=torch.int32, device=dev)   # rows >= n: reduce-scatter padding
    scores = padded[:n]
    mine = torch.empty((blk, n), dtype=torch.int32, device=dev) if world > 1 else None

    def step():
        padded.fill_(INT32_MIN)           # n*n*4 B write (> L2 for the N=1 workload): also the L2 flush
        eng.score(scores=scores)
        if world > 1:                     # rank r keeps rows [r0, r1) of the combined matrix
            dist.reduce_scatter_tensor(mine, padded, op=dist.ReduceOp.MAX)

    if world > 1:                         # untimed: the row block equals the all-reduced matrix's rows
        step()
        full = scores.clone()
        dist.all_reduce(full, op=dist.ReduceOp.MAX)
        assert bool((full[r0:r1] == mine[: r1 - r0]).all()), "reduce-scatter block differs from all-reduce"
        assert bool((full == full.T).all()) and bool((full.diagonal() > 0).all())
        del full

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dp_ms, launches = [], 0
    ev0.record()
    for _ in range(args.steps):
        step()
        launches += eng.last_launches()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dp_ms.append(eng.last_dp_ms())        # DP kernels of the last step, events on the launch stream
    clocks = sampler.stop() if rank == 0 else {}
    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(elapsed_ms.item()) / args.steps
    total_cells, total_pairs = plan["total_cells"], plan["total_pairs"]
    gcups = total_cells / (ms_per_step * 1e-3) / 1e9
    pdist = dist if world > 1 else None

    # ---- parity sample of the matrix leg: this rank's rows of the result just timed -------------
    parity = {"pairs": 0, "mismatches": 0}
    if not args.no_parity_sample:
        rng = np.random.default_rng(1000 + rank)
        m, bad = matrix_parity_sample(col, scores if world == 1 else mine[: r1 - r0], 0 if world == 1 else r0, rng)
        m, bad = _all_sum(pdist, dev, m, bad)
        parity["matrix_leg"] = {"pairs": m, "mismatches": bad,
                                "what": "random (i, j) of each rank's rows of the timed result vs oracle.pair_list"}
        parity["pairs"] += m
        parity["mismatches"] += bad

    # ---- e2e: host buffers in, host matrix out, copies inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        ascii_pin = torch.from_numpy(col.ascii.copy()).pin_memory()
        out_pin = torch.empty((n, n) if world == 1 else (blk, n), dtype=torch.int32).pin_memory()
        offsets = col.offsets

        def e2e_step():
            if world == 1:
                eng.score_all_host(ascii_pin.numpy(), offsets, out_pin.numpy())   # the C-ABI host call
            else:
                eng.pack_ascii(ascii_pin.numpy(), offsets)                        # H2D inside
                eng.plan(same_type_only=False, rank=rank, world=world)
                padded.fill_(INT32_MIN)
                eng.score(scores=scores)
                dist.reduce_scatter_tensor(mine, padded, op=dist.ReduceOp.MAX)
                out_pin.copy_(mine, non_blocking=False)                           # D2H inside, every rank its rows
        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reps = max(1, min(2, args.steps))
        for _ in range(reps):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": total_cells / float(t.item()) / 1e9, "unit": "GCUPS",
               "h2d_bytes_per_step": int(col.ascii.nbytes + col.offsets.nbytes + 20 * n) * world,
               "d2h_bytes_per_step": int(4 * n * n) if world == 1 else int(4 * blk * world * n),
               "ms_per_step": 1e3 * float(t.item()),
               "api": "bgca_score_all_host (C ABI, host pointers)" if world == 1 else
                      "Engine.pack_ascii+plan+score + NCCL reduce-scatter(max) + D2H of each rank's row block"}
        if rank == 0 and world == 1:
            m = out_pin.numpy()
            assert (np.diag(m) > 0).all() and (m[0] == m[:, 0]).all()

    # ---- the HBM-bound stages, measured live (rank 0, N=1 only; not part of the timed step) -------
    stages = None
    if rank == 0 and world == 1:
        try:
            hbm_peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            hbm_peak, peak_src = 6462.0, "fallback: copy bandwidth measured on this pool's B200s (B200_PROFILING.md)"
        from bgclib_b200 import synth as _synth
        big = _synth.nrps_pks_100k(4167)                       # BASELINE config 5: ~1e5 domains, 4167 BGCs
        nb, Bb = len(big), big.n_bgc
        ascii_dev = torch.from_numpy(big.ascii.copy()).to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        k1, k5 = [], []
        rowbest = torch.randint(0, 2000, (nb, Bb), dtype=torch.int32, device=dev)
        selfsc = torch.randint(1, 3000, (nb,), dtype=torch.int32, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(11):
            flush.fill_(1)                                     # evict L2 (126 MB) between repetitions
            eng.pack_ascii(ascii_dev, big.offsets, big.type_of_seq, big.bgc_of_seq, Bb)
            k1.append(eng.last_pack_ms())
            flush.fill_(2)
            e0.record()
            eng.reduce_bgc(rowbest, selfsc)
            e1.record()
            torch.cuda.synchronize()
            k5.append(e0.elapsed_time(e1))
        # best of 10, the way MEASURED_PEAKS.json took the copy figure these are divided by
        k1_ms, k5_ms = min(k1[1:]), min(k5[1:])
        k1_bytes = int(big.ascii.nbytes) + int(eng._lib.bgca_packed_bytes(eng._h))
        # rowbest read once, best written once, sim written once (the sim kernel's
        # reads of best / its transpose are re-reads of data just written, not algorithmic bytes)
        k5_bytes = 4 * nb * Bb + 8 * Bb * Bb + 8 * Bb * Bb
        stages = {"workload": f"{nb} synthetic NRPS/PKS domains in {Bb} BGCs (BASELINE config 5), cold L2",
                  "peak_gbs": hbm_peak, "peak_source": peak_src,
                  "timing": "CUDA events, L2 flushed before every repetition, best of 10 (as the peak was taken); medians in *_ms_median",
                  "k1_ms_median": statistics.median(k1[1:]), "k5_ms_median": statistics.median(k5[1:]),
                  "k1_pack": {"ms": k1_ms, "algorithmic_bytes": k1_bytes, "gbs": k1_bytes / k1_ms / 1e6,
                              "frac": k1_bytes / k1_ms / 1e6 / hbm_peak},
                  "k5_bgc_reduce": {"ms": k5_ms, "algorithmic_bytes": k5_bytes, "gbs": k5_bytes / k5_ms / 1e6,
                                    "frac": k5_bytes / k5_ms / 1e6 / hbm_peak}}
        del rowbest, selfsc, flush, ascii_dev
        eng.pack_ascii(col.ascii, col.offsets)                 # leave the engine on the bench workload

    # ---- fused BGC similarity, strong scaling on one fixed collection (configs[3]; [4] at N = 8) ----
    fused = fused5 = api_e2e = None
    if not args.no_fused:
        from bgclib_b200 import synth as _synth
        col4 = _synth.mibig_scale(2500)
        frng = None if args.no_parity_sample else np.random.default_rng(2000 + rank)
        fused, chk = fused_leg(eng, col4, pdist, rank, world, dev, max(1, args.fused_steps), 1, True, frng)
        if chk is not None:
            npairs, m, bad = _all_sum(pdist, dev, *chk)
            parity["fused_rowbest"] = {"entries": m, "pairs": npairs, "mismatches": bad,
                                       "what": "random rowbest[d, b] of the combined fused result vs the max of "
                                               "oracle.pair_list over the same-type domains of BGC b"}
            parity["pairs"] += npairs
            parity["mismatches"] += bad
            ent, bad = subcollection_parity(eng, col4, pdist, rank, world, dev)
            ent, bad = _all_sum(pdist, dev, ent, bad)
            parity["fused_subcollection"] = {"entries": ent, "mismatches": bad,
                                             "what": "best, self, sim of the first 300 domains (all ranks) vs "
                                                     "oracle.all_pairs + oracle.bgc_similarity_np on rank 0"}
            parity["pairs"] += 300 * 301 // 2
            parity["mismatches"] += bad
        # ---- the same collection as OBJECTS through the public Python API (N = 1): how much of
        # bgc_similarity(collection) is spent outside the DP kernels (extraction + pack) ----------
        if world == 1:
            import time as _t
            from bgclib_b200 import api as _api, extract as _extract
            objs = _synth.as_collection(col4)
            t0 = _t.perf_counter()
            batch = _extract(objs)
            t_extract = _t.perf_counter() - t0
            torch.cuda.synchronize()
            t0 = _t.perf_counter()
            eng.pack_batch(batch, type_of_seq=batch.type_of_seq, bgc_of_seq=batch.bgc_of_seq, n_bgc=len(batch.bgc_ids))
            torch.cuda.synchronize()
            t_pack = _t.perf_counter() - t0
            t0 = _t.perf_counter()
            res = _api.bgc_similarity(objs, device=local_rank)          # host objects in, host matrices out
            torch.cuda.synchronize()
            t_call = _t.perf_counter() - t0
            api_e2e = {"api": "bgclib_b200.bgc_similarity(collection of BGC/protein/domain objects) -> numpy sim, best, self",
                       "workload": f"{len(batch)} domains on {sum(len(g.protein_list) for g in objs.bgcs.values())} "
                                   f"proteins in {len(objs.bgcs)} BGCs (configs[3] as objects)",
                       "call_s": t_call, "gcups": fused["cells"] / t_call / 1e9,
                       "extract_s": t_extract, "pack_s": t_pack, "extract_plus_pack_share": (t_extract + t_pack) / t_call,
                       "residue_buffer_bytes": int(batch.spans()[0].nbytes),
                       "sim_shape": list(res.sim.shape), "sim_diag_ones": bool((np.diag(res.sim) == 1).all())}
            del objs, batch, res
        del col4
        if args.fused_config5 or world >= 8:
            col5 = _synth.nrps_pks_100k(4167)
            fused5, _ = fused_leg(eng, col5, pdist, rank, world, dev, 1, 0, False, None)
            del col5
        eng.pack_ascii(col.ascii, col.offsets)                 # leave the engine on the bench workload

    if rank == 0:
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz", clocks.get("sm_max_mhz") or 1965.0))
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        # Peak = the sm_100a issue-rate limit north_star names: one warp instruction per clock per
        # scheduler = 128 lanes/clk/SM, at the maximum SM clock.  (The alu pipe alone takes half of
        # that; the kernel puts its packed adds on the fma pipe, so the alu-pipe nominal is not a
        # ceiling for it — kept beside the fraction as frac_alu_pipe_nominal.)
        peak_issue = n_sm * ISSUE_LANES_PER_CLK_SM * sm_max * 1e6 / 1e12        # Tlane-op/s
        peak_alu = n_sm * ALU_LANES_PER_CLK_SM * sm_max * 1e6 / 1e12
        achieved = plan["cells"] * ALG_LANE_OPS_PER_CELL / (dp_ms[-1] * 1e-3) / 1e12
        roofline = {
            "bound": "issue", "kernel": "gotoh_pair16_kernel<G,K,local> (all variants of one step)",
            "achieved": achieved, "peak": peak_issue, "unit": "Tlane-op/s", "frac": achieved / peak_issue,
            "peak_source": f"sm_100a issue rate: {n_sm} SMs x {ISSUE_LANES_PER_CLK_SM} lanes/clk x {sm_max:.0f} MHz "
                           "(MEASURED_PEAKS.json holds no integer/DPX figure; B300_MICROARCH.md 'Instruction issue'; "
                           "dpx_probe below measures the per-op rates)",
            "alg_lane_ops_per_cell": ALG_LANE_OPS_PER_CELL,
            "kernel_ms": dp_ms[-1], "kernel_gcups": plan["cells"] / (dp_ms[-1] * 1e-3) / 1e9,
            "frac_at_observed_clock": achieved / (n_sm * ISSUE_LANES_PER_CLK_SM * sm_mhz * 1e6 / 1e12),
            "frac_alu_pipe_nominal": achieved / peak_alu,
            "traffic": None,
        }
        try:
            # dram__bytes_read.sum + dram__bytes_write.sum of one DP launch from the committed
            # `ncu --set full` capture, scaled by launch time to the launches of one step (the DP
            # launches differ only in K; their DRAM bytes per ms are the same within 10 %)
            if world > 1:
                # the committed capture is a 1-GPU step; under N ranks every rank fills, reads and permutes
                # the whole staging matrix for 1/N of the pairs, which this scaling does not describe
                raise LookupError("traffic is reported for the 1-GPU step only")
            tr = json.loads((ROOT / "profiles" / "dp_traffic.json").read_text())
            entries = float(n) * float(n)
            dp_part = tr["dram_bytes_per_launch_ms"] * dp_ms[-1]
            perm_part = tr["permute_dram_bytes_per_entry"] * entries
            wb_part = tr["s16_writeback_bytes_per_entry"] * entries
            alg = 8 * plan["n_pairs"] + int(col.ascii.nbytes)
            roofline["traffic"] = dp_part + perm_part + wb_part
            roofline["traffic_over_algorithmic"] = roofline["traffic"] / alg
            roofline["traffic_note"] = (
                f"DRAM bytes per step from the committed ncu capture ({tr['source']}): DP launches "
                f"{tr['dram_bytes_per_launch']} B per {tr['launch_ms']:.1f} ms launch, scaled by DP time = {dp_part:.3g} B; "
                f"+ the sorted-space int16 staging matrix written back from L2 at most once = {wb_part:.3g} B (upper bound, ncu "
                f"attributes it to no launch); + permute16_kernel {perm_part:.3g} B (reads the staging matrix, writes the "
                f"caller's int32 matrix); algorithmic bytes per step = {alg} (two int32 stores per pair + residues); "
                "round 1 stored scores straight into the caller's index space: 2.73e9 B per step")
        except Exception:
            pass
        if (args.probe or world == 1) and not args.no_probe:
            probe = bgclib_b200.dpx_probe(local_rank)
            roofline["dpx_probe_lane_ops_per_clk_sm"] = probe
            # the bare DP instruction stream of this kernel (no loads, shuffles, loop): the measured
            # issue-rate ceiling for this op mix, as GCUPS at the clock observed during the run
            stream_gcups = n_sm * (probe["cell_mix_s16x2"] / 5.5) * 2 * sm_mhz * 1e6 / 1e9
            roofline["peak_measured_dp_stream_gcups"] = stream_gcups
            roofline["frac_of_measured_dp_stream"] = roofline["kernel_gcups"] / stream_gcups
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            g, pps, cores, desc, _ = cpu_reference_run(col, args.cpu_seconds)
            cpu = {"value": g, "unit": "GCUPS", "cores": cores, "kind": "port", "sample": desc,
                   "pairs_per_s": pps}
        line = {
            "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "s16x2", "data": "synthetic",
            "config": bench_config(wl),
            "pairs": total_pairs, "cells": total_cells, "parallelism": f"tile-sharded x{world}",
            "l2": "each step rewrites the n*n int32 score matrix (400 MB at N=1) before the kernels, which "
                  "evicts L2; DP inputs (4 MB of residues) are L2-resident by design",
            "tiles_rank0": plan["n_tiles"], "kernel_variants_rank0": plan["n_variants"],
            "pairs_per_s": total_pairs / (ms_per_step * 1e-3),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu, "hbm_stages": stages,
            "fused": fused, "fused_config5": fused5, "api_e2e": api_e2e, "parity_sample": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
