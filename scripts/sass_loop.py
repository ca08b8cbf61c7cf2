#!/usr/bin/env python
"""Opcode histogram of the innermost loops of a kernel that contain TMEM loads (the tCG loops of the Sphere kernels).

    python scripts/sass_loop.py <cubin or .so> <kernel name substring> [--dump FILE]
"""
import collections
import re
import subprocess
import sys

path, name = sys.argv[1], sys.argv[2]
dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
ins, on = [], False
for line in sass.splitlines():
    if "Function :" in line:
        on = name in line
        if on:
            print(line.strip())
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
idx = {a: i for i, (a, _) in enumerate(ins)}
print("instructions:", len(ins))
loops = []
for i, (a, t) in enumerate(ins):
    if "BRA" in t:
        m = re.search(r"0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in idx:
            j = idx[int(m.group(1), 16)]
            body = ins[j:i + 1]
            nld = sum("LDTM" in x for _, x in body)
            if nld:
                loops.append((len(body), nld, j, i))
loops.sort()
for n, nld, j, i in loops[:2]:
    c = collections.Counter()
    for a, t in ins[j:i + 1]:
        f = t.split()
        op = f[1] if f[0].startswith("@") else f[0]
        c[op.split(".")[0]] += 1
    print(f"loop {ins[j][0]:#x}..{ins[i][0]:#x}: {n} instructions, {nld} LDTM:", dict(c.most_common()))
    if dump:
        with open(dump, "w") as f:
            f.write("\n".join(f"{a:05x} {t}" for a, t in ins[j:i + 1]))
        dump = None
