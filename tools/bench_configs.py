#!/usr/bin/env python
"""Stage-level measurements on the BASELINE.json configs that bench.py does not
headline (configs 4/5 and the HBM-bound stages K1/K5).  One JSON line per
measurement; run under torchrun for N > 1.

    python tools/bench_configs.py --config 4 [--same-type] [--matrix] [--steps 2]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4, choices=[3, 4, 5])
    ap.add_argument("--n-bgc", type=int, default=None)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--same-type", action="store_true", help="only same-type pairs (the BGC similarity plan)")
    ap.add_argument("--matrix", action="store_true", help="also materialise the n x n score matrix")
    ap.add_argument("--mode", default="local")
    ap.add_argument("--stages-only", action="store_true",
                    help="K1 and K5 alone (rowbest filled with random scores instead of running the DP kernels)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import bgclib_b200
    from bgclib_b200 import synth
    from bgclib_b200.api import INT32_MIN

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t0 = time.perf_counter()
    if args.config == 3:
        col = synth.ks_like(10_000)
    elif args.config == 4:
        col = synth.mibig_scale(args.n_bgc or 2500)
    else:
        col = synth.nrps_pks_100k(args.n_bgc or 4167)
    gen_s = time.perf_counter() - t0
    n, B = len(col), col.n_bgc
    eng = bgclib_b200.get_engine(local_rank)
    eng.set_scoring("BLOSUM62", -12, -1, args.mode)

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    def sync_time(fn, reps=1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- K1 packer: ASCII already on the device, HBM-bound ---------------------------------
    ascii_dev = torch.from_numpy(col.ascii.copy()).to(dev)
    types = col.type_of_seq
    eng.pack_ascii(ascii_dev, col.offsets, types, col.bgc_of_seq, B)       # warm-up (allocations)
    t_pack = time.perf_counter()
    eng.pack_ascii(ascii_dev, col.offsets, types, col.bgc_of_seq, B)
    torch.cuda.synchronize()
    pack_wall_ms = 1e3 * (time.perf_counter() - t_pack)
    packed_bytes = int(eng._lib.bgca_packed_bytes(eng._h))
    emit({"stage": "K1 pack (host sort + uploads + kernel, wall)", "config": args.config, "n": n,
          "ascii_bytes": int(col.ascii.nbytes), "packed_bytes": packed_bytes, "wall_ms": pack_wall_ms,
          "note": "kernel alone is in the ncu launch list; algorithmic bytes = ascii + packed"})

    # ---- plan ---------------------------------------------------------------------------------
    t_plan = time.perf_counter()
    plan = eng.plan(same_type_only=args.same_type, rank=rank, world=world)
    plan_ms = 1e3 * (time.perf_counter() - t_plan)
    emit({"stage": "K6 plan", "config": args.config, "same_type_only": args.same_type, "wall_ms": plan_ms, **plan})

    # ---- fused path: DP + rowbest epilogue, then K5 ---------------------------------------------
    rowbest = torch.zeros((n, B), dtype=torch.int32, device=dev)
    selfsc = torch.zeros((n,), dtype=torch.int32, device=dev)

    def fused():
        rowbest.zero_()
        selfsc.zero_()
        eng.score(rowbest=rowbest, selfsc=selfsc, type_check=False)
        if world > 1:
            dist.all_reduce(rowbest, op=dist.ReduceOp.MAX)
            dist.all_reduce(selfsc, op=dist.ReduceOp.MAX)

    if args.stages_only:
        rowbest.random_(0, 2000)
        selfsc.random_(1, 3000)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2: evicts between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        alg5 = 4 * n * B + 4 * n + 8 * B * B + 8 * B + (8 * B * B * 2 + 8 * B * B)
        alg1 = int(col.ascii.nbytes) + packed_bytes
        t1, t5 = [], []
        for _ in range(6):
            flush.fill_(1)
            eng.pack_ascii(ascii_dev, col.offsets, types, col.bgc_of_seq, B)   # K1 (kernel time: ncu launch list)
            flush.fill_(2)
            e0.record()
            eng.reduce_bgc(rowbest, selfsc)
            e1.record()
            torch.cuda.synchronize()
            t5.append(e0.elapsed_time(e1))
        ms5 = float(np.median(t5[1:]))
        emit({"stage": "K5 bgc_sum + bgc_sim (cold L2, CUDA events)", "config": args.config, "n": n, "n_bgc": B,
              "ms": ms5, "algorithmic_bytes": alg5, "gb_per_s": alg5 / (ms5 * 1e-3) / 1e9})
        emit({"stage": "K1 pack_kernel", "config": args.config, "n": n, "algorithmic_bytes": alg1,
              "note": "kernel duration: gpu__time_duration of pack_kernel in the ncu launch list of this command"})
        emit({"stage": "done", "generate_s": gen_s})
        return
    if args.mode == "local":
        fused()
        ms = sync_time(fused, args.steps)
        emit({"stage": "K2/K3 + fused K5 epilogue (rowbest atomics) + NCCL combine", "config": args.config,
              "n_gpus": world, "n": n, "n_bgc": B, "same_type_only": args.same_type, "ms": ms,
              "gcups": plan["total_cells"] / (ms * 1e-3) / 1e9, "pairs_per_s": plan["total_pairs"] / (ms * 1e-3),
              "dp_ms_rank0": eng.last_dp_ms(), "launches": eng.last_launches(), "cells": plan["total_cells"],
              "pairs": plan["total_pairs"]})
        out = {}

        def k5():
            out["r"] = eng.reduce_bgc(rowbest, selfsc)
        k5()
        ms5 = sync_time(k5, 5)
        alg = 4 * n * B + 4 * n + 8 * B * B + 8 * B + (8 * B * B * 2 + 8 * B * B)   # sum kernel + sim kernel
        emit({"stage": "K5 bgc_sum + bgc_sim", "config": args.config, "n": n, "n_bgc": B, "ms": ms5,
              "algorithmic_bytes": alg, "gb_per_s": alg / (ms5 * 1e-3) / 1e9})
        sim = out["r"][0]
        emit({"stage": "check", "sim_diag_ones": bool((torch.diagonal(sim) == 1).all().item()),
              "sim_max": float(sim.max().item()), "sim_symmetric": bool((sim == sim.T).all().item())})

    # ---- matrix path --------------------------------------------------------------------------
    if args.matrix:
        scores = torch.empty((n, n), dtype=torch.int32, device=dev)

        def mat():
            scores.fill_(INT32_MIN)
            eng.score(scores=scores)
            if world > 1:
                dist.all_reduce(scores, op=dist.ReduceOp.MAX)
        mat()
        ms = sync_time(mat, args.steps)
        emit({"stage": "K2/K3 -> n x n int32 matrix + NCCL combine", "config": args.config, "n_gpus": world,
              "n": n, "same_type_only": args.same_type, "mode": args.mode, "ms": ms,
              "gcups": plan["total_cells"] / (ms * 1e-3) / 1e9, "pairs_per_s": plan["total_pairs"] / (ms * 1e-3),
              "dp_ms_rank0": eng.last_dp_ms(), "launches": eng.last_launches()})
    emit({"stage": "done", "generate_s": gen_s})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
