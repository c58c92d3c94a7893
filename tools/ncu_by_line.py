#!/usr/bin/env python
"""Attribute ncu warp-stall samples of voxel_pipeline_kernel<true> to CUDA source lines.

usage: ncu_by_line.py report.ncu-rep libdecaes_cuda.so [topN] [file-filter]
Joins the SASS page of the report with `nvdisasm -g` line info of the cubin embedded in the .so
(built with -lineinfo) by instruction offset.
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, so = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
filt = sys.argv[4] if len(sys.argv) > 4 else ""
kern = os.environ.get("KERN", "voxel_pipeline_kernelILb1ELb0E")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}
cur = None
insec = False
for ln in dis.splitlines():
    if ln.startswith("//---------------------"):
        insec = (".text." in ln and kern in ln)
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:voxel_pipeline"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    off = int(r[0], 16) - base
    key = line_of.get(off, ("?", 0))
    a = agg[key]
    a["samples"] += int(r[ix["# Samples"]])
    a["inst"] += int(r[ix["Instructions Executed"]])
    for s in stalls:
        a[s] += int(r[ix[s]])
tot = sum(a["samples"] for a in agg.values())
src_cache = {}
def src(f, n):
    for d in ("decaes.jl_b200/csrc", "include"):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
    return ""
items = [(k, a) for k, a in agg.items() if filt in k[0]]
for (f, n), a in sorted(items, key=lambda kv: -kv[1]["samples"])[:topn]:
    top = sorted(((a[s], s[6:]) for s in stalls), reverse=True)[:3]
    tops = " ".join(f"{nm}:{100*c/max(a['samples'],1):.0f}" for c, nm in top)
    print(f"{100*a['samples']/tot:5.2f}% {a['inst']/1e6:8.1f}Mi {f}:{n:<5d} {tops:40s} | {src(f, n)}")
