#!/usr/bin/env python
"""Opcode histogram of a kernel's hot loop from `cuobjdump -sass` of the built library — the
committed evidence that the DP kernels are made of sm_100a DPX instructions.

    python tools/sass_hot_loop.py > profiles/r02_sass_hot_loop.txt

For every kernel listed below: the function's SASS is cut at its longest branch-free stretch
that contains DPX / FP64-compare opcodes (the unrolled DP step), and the opcodes in it counted.
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "bgclib_b200" / "libbgca.so"
KERNELS = [
    ("gotoh_pair16_kernel<16, 28, 0, 1, 0, 0>", "hot kernel of the bench: G=16 lanes x K=28 rows, local, gap preset -12/-1"),
    ("gotoh_pair16_kernel<16, 28, 1, 1, 0, 0>", "the same, global mode"),
    ("gotoh_pair16_kernel<1, 20, 0, 1, 0, 0>", "one lane per alignment (20-residue domains)"),
    ("gotoh_pair16_kernel<32, 24, 0, 1, 1, 0>", "multi-pass whole-warp form (queries > 1024 residues)"),
    ("gotoh_s32_kernel<0>", "K4: 32-bit fallback, local"),
    ("gotoh_stats2_kernel<0, 14, 0>", "K7: statistics, FP64-compare form, 14 rows per lane, running best in the loop"),
    ("gotoh_stats2_kernel<0, 14, 1>", "K7 with score hints"),
]
INTERESTING = re.compile(r"VIMNMX|VIADDMNMX|VIADD|DSETP")


def function_sass(name):
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    chunks = out.split("Function : ")
    res = []
    for c in chunks[1:]:
        mangled = c.split("\n", 1)[0].strip()
        dem = subprocess.run(["cu++filt", mangled], capture_output=True, text=True).stdout.strip()
        if name in dem.replace("bgca::", "").replace("(int)", ""):
            res.append((dem, c))
    return res


def opcodes(text):
    ops = []
    for line in text.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1).strip()
        if ins.startswith("@"):
            ins = ins.split(None, 1)[1]
        ops.append(ins.split()[0])
    return ops


def hot_stretch(ops):
    best, cur = [], []
    for op in ops + ["BRA"]:
        if op.split(".")[0] in ("BRA", "EXIT", "RET", "BSSY", "BSYNC", "WARPSYNC", "CALL"):
            if sum(1 for o in cur if INTERESTING.match(o)) > sum(1 for o in best if INTERESTING.match(o)):
                best = cur
            cur = []
        else:
            cur.append(op)
    return best


def step_census(text):
    """Instructions one step of the streaming loop executes on its hot path: the innermost loop
    (backward branch) that holds the DP instructions, walked from its head with unconditional
    forward branches followed (they jump over the cold boundary / idle code) and conditional ones
    not taken."""
    ins = []
    for line in text.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    pos = {a: i for i, (a, _) in enumerate(ins)}
    def walk(first, last):
        hot, i = [], first
        while i <= last:
            a, t = ins[i]
            hot.append(t.split(None, 1)[1].split()[0] if t.startswith("@") else t.split()[0])
            m = re.match(r"BRA\s+(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt > a and tgt in pos:
                    i = pos[tgt]
                    continue
            i += 1
        return len(hot), sum(1 for o in hot if INTERESTING.match(o))

    # where the unrolled DP step sits: the longest branch-free run with DP instructions
    def opname(t):
        return t.split(None, 1)[1].split()[0] if t.startswith("@") else t.split()[0]
    run0, cur0, stretch = 0, 0, (0, 0, -1)
    for i, (_, t) in enumerate(ins + [(0, "BRA")]):
        if opname(t).split(".")[0] in ("BRA", "EXIT", "RET", "BSSY", "BSYNC", "WARPSYNC", "CALL"):
            n = sum(1 for _, x in ins[cur0:i] if INTERESTING.match(opname(x)))
            if n > stretch[2]:
                stretch = (cur0, i - 1, n)
            cur0 = i + 1
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(0x[0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in pos and pos[tgt] <= stretch[0] and i >= stretch[1]:
            if best is None or i - pos[tgt] < best[1] - best[0]:      # the innermost loop around it
                best = (pos[tgt], i)
    return walk(*best) if best else None


def main():
    print(f"# cuobjdump -sass {LIB.relative_to(ROOT)}  (sm_100a); tools/sass_hot_loop.py")
    for name, what in KERNELS:
        found = function_sass(name)
        if not found:
            print(f"\n## {name}: NOT FOUND")
            continue
        dem, text = found[0]
        ops = opcodes(text)
        hot = hot_stretch(ops)
        c = Counter(o.replace(".reuse", "") for o in hot)
        print(f"\n## {name} — {what}")
        print(f"#  whole function: {len(ops)} instructions; hot stretch (longest branch-free run with DPX/DSETP): {len(hot)}")
        sc = step_census(text)
        if sc:
            print(f"#  one step of the streaming loop, hot path: {sc[0]} instructions, {sc[1]} of them DPX / packed add / DSETP")
        for op, k in c.most_common():
            print(f"{k:6d}  {op}")


if __name__ == "__main__":
    sys.exit(main())
