"""In-tree build of libbgca.so (hand-written sm_100a CUDA + the C ABI).

    python -m bgclib_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is linked with a static cudart so it
loads on a CPU-only box (the planner entry points work there; everything else
reports BGCA_ERR_CUDA).  Objects go to bgclib_b200/csrc/_build/, the library to
bgclib_b200/libbgca.so (git-ignored; travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = CSRC / "_build"
LIB = PKG / "libbgca.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--cudart", "static",
                 "-diag-suppress", "177"]

# (object name, source, extra defines)
UNITS = [("bgca_api.o", "bgca_api.cu", []),
         ("pack_reduce.o", "pack_reduce.cu", []),
         ("stats.o", "stats.cu", []),
         ("stats2.o", "stats2.cu", []),
         ("dpx_probe.o", "dpx_probe.cu", [])]
for g in (1, 2, 4, 8, 16, 32):
    for m in (0, 1):
        for pr in (0, 1, 2):
            UNITS.append((f"dp_g{g}_m{m}_p{pr}.o", "dp_inst.cu",
                          [f"-DBGCA_INST_G={g}", f"-DBGCA_INST_MODE={m}", f"-DBGCA_INST_PRESET={pr}"]))


for m in (0, 1):                 # multi-pass whole-warp instantiations (queries > 1024 residues), and
    for pr in (0, 1, 2):         # their 32-bit form (one alignment per word: scores beyond int16)
        for w32 in (0, 1):
            UNITS.append((f"dp_mp{'w' if w32 else ''}_m{m}_p{pr}.o", "dp_inst.cu",
                          ["-DBGCA_INST_G=32", "-DBGCA_INST_MP=1", f"-DBGCA_INST_W32={w32}",
                           f"-DBGCA_INST_MODE={m}", f"-DBGCA_INST_PRESET={pr}"]))


def _sources_mtime() -> float:
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    files.append(PKG.parent / "include" / "bgca.h")
    return max(f.stat().st_mtime for f in files)


def is_stale() -> bool:
    return (not LIB.exists()) or LIB.stat().st_mtime < _sources_mtime()


def _compile(unit):
    obj, src, defs = unit
    cmd = [NVCC] + CFLAGS + defs + ["-c", str(CSRC / src), "-o", str(OBJ / obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src} {defs}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    OBJ.mkdir(exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(_compile, UNITS))
    cmd = [NVCC] + ARCH + ["-shared", "--cudart", "static", "-o", str(LIB)] + [str(OBJ / o) for o in objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
