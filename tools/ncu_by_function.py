#!/usr/bin/env python
"""Attribute ncu warp-stall samples of voxel_pipeline_kernel to its out-of-line device functions.

usage: ncu_by_function.py report.ncu-rep libdecaes_cuda.so [kernel-substring]
Reads the SASS source page of the report, the symbol table of the cubin embedded in the .so
(`cuobjdump -elf`), and prints samples / instructions / top stall reasons per function.
"""
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
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
def owner(off):
    for a, s, n in syms:
        if a <= off < a + s:
            return n
    return "<kernel body>"
for r in rows[2:]:
    off = int(r[0], 16) - base
    f = owner(off)
    agg[f]["samples"] += int(r[ix["# Samples"]])
    agg[f]["inst"] += int(r[ix["Instructions Executed"]])
    agg[f]["sass"] += 1
    for s in stalls:
        agg[f][s] += int(r[ix[s]])
tot = sum(a["samples"] for a in agg.values())
toti = sum(a["inst"] for a in agg.values())
print(f"total samples {tot}, warp instructions {toti}")
print(f"{'function':58s} {'samp%':>6s} {'inst%':>6s} {'SASS':>6s}  top stalls")
for f, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    top = sorted(((a[s], s[6:]) for s in stalls), reverse=True)[:4]
    tops = " ".join(f"{n}:{100*c/max(a['samples'],1):.0f}" for c, n in top)
    print(f"{f[:58]:58s} {100*a['samples']/tot:6.2f} {100*a['inst']/toti:6.2f} {a['sass']:6d}  {tops}")
