#!/usr/bin/env python
"""Per-callee instruction / sample shares of one kernel from an `ncu --page source --csv` export (gzip): the kernel body up
to the first out-of-line device function, then each CALL.REL.NOINC target.  Usage: ncu_callees.py <source.csv.gz>"""
import collections, csv, gzip, io, re, sys

def load(path):
    rows = list(csv.reader(io.StringIO(gzip.open(path, "rt").read())))
    for i, r in enumerate(rows):
        if "Source" in r and any("Instructions Executed" in c for c in r):
            return rows[i], rows[i + 1:]
    raise SystemExit("no source table")

hdr, rows = load(sys.argv[1])
iA, iS, iE, iSm = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = []
for r in rows:
    try:
        ins.append((int(r[iA], 16) if r[iA].startswith("0x") else int(r[iA]), r[iS], int(r[iE]), int(r[iSm])))
    except Exception:
        pass
base = ins[0][0]
tot, tots = sum(x[2] for x in ins), sum(x[3] for x in ins)
byop, sampop = collections.Counter(), collections.Counter()
for a, s_, e, sm in ins:
    t = s_.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op] += e
    sampop[op] += sm
print(f"total warp instructions {tot:.4e}, samples {tots}")
print("  " + "  ".join(f"{op} {100 * c / tot:.1f}%/{100 * sampop[op] / tots:.1f}%" for op, c in byop.most_common(14)))
targets = collections.Counter()
for a, s_, e, sm in ins:
    m = re.search(r"CALL\.REL\.NOINC\s+(0x[0-9a-f]+)", s_)
    if m:
        targets[int(m.group(1), 16)] += e
entries = sorted(targets)
bounds = entries + [ins[-1][0] + 16]
first = entries[0] if entries else ins[-1][0] + 16
print(f"kernel body: {100 * sum(x[2] for x in ins if x[0] < first) / tot:.1f}% instr, "
      f"{100 * sum(x[3] for x in ins if x[0] < first) / tots:.1f}% samples")
for k, ent in enumerate(entries):
    reg = [x for x in ins if ent <= x[0] < bounds[k + 1]]
    e, sm = sum(x[2] for x in reg), sum(x[3] for x in reg)
    ops = collections.Counter()
    for x in reg:
        t = x[1].split()
        ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += 1
    print(f"callee @{ent - base:x}: calls {targets[ent]:.3e} static {len(reg)} instr {100 * e / tot:.1f}% ({e / max(targets[ent], 1):.0f}/call) "
          f"samples {100 * sm / tots:.1f}%  {dict(ops.most_common(5))}")
