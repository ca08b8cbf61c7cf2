#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of csrc/libriptrm_b200.so (cuobjdump -sass), written to profiles/r02_sass_summary.txt:
the evidence table of B200_PROFILING.md "What proves a Blackwell-native kernel" (LDTM / STTM / UTCATOMSWS = tcgen05 Tensor
Memory traffic and allocation, UBLKCP / SYNCS = bulk TMA copies and mbarrier waits, DMMA = the FP64 tensor path -- tcgen05.mma has
no f64 kind --, DFMA = the FP64 vector pipe).

    python scripts/sass_summary.py [--out profiles/r02_sass_summary.txt]
"""
import argparse
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "riemannian-interior-point-trust-region-method_b200", "csrc", "libriptrm_b200.so")
COLS = ("LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "MUFU", "SHFL", "LDS", "STS",
        "LDG", "STG", "BAR", "ATOM")


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def short(name):
    if name.endswith(")"):                                 # drop the parameter list (the last balanced parenthesis group)
        depth = 0
        for i in range(len(name) - 1, -1, -1):
            depth += (name[i] == ")") - (name[i] == "(")
            if depth == 0:
                name = name[:i]
                break
    name = re.sub(r"\((?:int|bool)\)", "", name)
    name = name.replace("riptrm::", "").replace("void ", "")
    return name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "r02_sass_summary.txt"))
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, total, cur = collections.OrderedDict(), {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            total[cur] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            op = m.group(1)
            total[cur] += 1
            for c in COLS:
                if op == c or op.startswith(c):
                    counts[cur][c] += 1
                    break
    names = demangle(list(counts))
    rows = sorted(counts, key=lambda k: -total[k])
    width = 64
    lines = [f"SASS mnemonic counts per kernel: {os.path.relpath(LIB, REPO)}  (cuobjdump -sass; arch {', '.join(arch)})",
             "static instruction counts (not executed counts); a kernel template instantiation per row", "",
             f"{'kernel':<{width}}  {'instr':>7}  " + "  ".join(f"{c:>6}" for c in COLS)]
    for k in rows:
        lines.append(f"{short(names[k])[:width]:<{width}}  {total[k]:>7}  " + "  ".join(f"{counts[k][c]:>6}" for c in COLS))
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))
    print(f"... {len(rows)} kernels -> {args.out}")


if __name__ == "__main__":
    main()
