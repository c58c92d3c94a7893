#!/usr/bin/env python
"""Instruction-cache footprint of voxel_pipeline_kernel from an ncu report: 128-byte lines (8 SASS instructions) ranked by
executions, per out-of-line function.  usage: ncu_hot_lines.py report.ncu-rep lib.so [kernel-substring]"""
import collections, csv, re, subprocess, sys
rep, so = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "voxel_pipeline_kernelILb1ELb0E"
elf = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
syms = []
for line in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", line)
    if m and kern in m.group(3):
        syms.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
syms = sorted(set(syms))
dem = subprocess.run(["c++filt"] + [s[2] for s in syms], capture_output=True, text=True).stdout.splitlines()
syms = [(a, s, re.sub(r"\(.*", "", d).replace("decaes::", "")) for (a, s, _), d in zip(syms, dem)]
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:voxel_pipeline"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
def owner(off):
    for a, s, n in syms:
        if a <= off < a + s:
            return n
    return "<kernel body>"
lines = collections.Counter()   # line index -> max executions of an instruction in the line ~ fetches of the line
lown = {}
for r in rows[2:]:
    off = int(r[0], 16) - base
    ex = int(r[ix["Instructions Executed"]])
    li = off // 128
    lines[li] = max(lines[li], ex)
    lown.setdefault(li, owner(off))
tot = sum(lines.values())
print(f"lines touched: {sum(1 for v in lines.values() if v)} of {len(lines)}; line fetch events (upper bound) {tot}")
acc = 0
for frac in (0.5, 0.8, 0.9, 0.95, 0.99):
    acc = 0
    n = 0
    for li, v in lines.most_common():
        acc += v
        n += 1
        if acc >= frac * tot:
            break
    print(f"  {int(frac*100)} % of the line executions come from {n} lines = {n * 128 // 1024} KB")
byf = collections.defaultdict(lambda: [0, 0, 0])
thr = 0.0005 * tot
for li, v in lines.items():
    f = lown[li]
    byf[f][0] += 1
    byf[f][1] += v
    if v >= thr:
        byf[f][2] += 1
print(f"{'function':58s} {'lines':>6s} {'hot':>5s} {'exec%':>7s}")
for f, (n, e, h) in sorted(byf.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{f[:58]:58s} {n:6d} {h:5d} {100 * e / tot:7.2f}")
