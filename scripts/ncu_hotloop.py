#!/usr/bin/env python
"""Where a kernel's hottest loop spends its stall samples, from an `ncu --page source --csv` export (gzip).

    python scripts/ncu_hotloop.py <source.csv.gz> [--launch K] [--list]

The hot loop = the instructions executed at least half as often as the most-executed one.  Prints the loop's share of all
samples, samples per stall reason, and samples per opcode; --list prints every instruction of the loop with its samples."""
import collections, csv, gzip, io, sys

rows = list(csv.reader(io.StringIO(gzip.open(sys.argv[1], "rt").read())))
# an export of several launches holds one table per launch ("Kernel Name" rows): --launch K picks one (default: the last)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] or [0]
which = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else len(starts) - 1
rows = rows[starts[which]:(starts[which + 1] if which + 1 < len(starts) else len(rows))]
for i, r in enumerate(rows):
    if "Source" in r and any("Instructions Executed" in c for c in r):
        hdr, rows = rows[i], rows[i + 1:]
        break
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(j, c) for j, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
ins = []
for r in rows:
    try:
        ins.append((r[iS].strip(), int(r[iE]), int(r[iSm]), [int(r[j] or 0) for j, _ in stall_cols]))
    except Exception:
        pass
mx = max(x[1] for x in ins)
hot = [k for k, x in enumerate(ins) if x[1] >= 0.5 * mx]
lo, hi = min(hot), max(hot)
loop = [x for x in ins[lo:hi + 1] if x[1] >= 0.5 * mx]
tot_s = sum(x[2] for x in ins)
ls = sum(x[2] for x in loop)
le = sum(x[1] for x in loop)
print(f"kernel: {len(ins)} instructions, {sum(x[1] for x in ins):.4e} executed, {tot_s} samples")
print(f"hot loop: {len(loop)} instructions, {100.0 * le / sum(x[1] for x in ins):.1f}% of executed, {100.0 * ls / tot_s:.1f}% of samples, "
      f"{ls / max(1, le) * mx:.0f} samples per iteration-slot")
reasons = collections.Counter()
for x in loop:
    for (j, c), v in zip(stall_cols, x[3]):
        reasons[c] += v
tr = sum(reasons.values())
print("stall reasons:", "  ".join(f"{c[6:]} {100.0 * v / tr:.1f}%" for c, v in reasons.most_common(9)))
byop, cnt = collections.Counter(), collections.Counter()
for x in loop:
    t = x[0].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op] += x[2]
    cnt[op] += 1
print("samples by opcode (count in loop):", "  ".join(f"{op} {100.0 * v / ls:.1f}% ({cnt[op]})" for op, v in byop.most_common(14)))
if "--list" in sys.argv:
    for x in loop:
        top = max(zip(x[3], [c for _, c in stall_cols]))
        print(f"{x[2]:7d} {100.0 * x[2] / ls:5.2f}%  {top[1][6:]:14s} {x[0]}")
