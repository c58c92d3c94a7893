#!/usr/bin/env python
"""Instruction census of the shipped kernels (cuobjdump -sass): tensor-core, TMA, REDUX, barrier and FP64 mnemonics per kernel."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "decaes.jl_b200/libdecaes_cuda.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["DMMA", "UBLKCP", "SYNCS", "REDUX", "CREDUX", "BAR", "DFMA", "DADD", "DMUL", "MUFU.RSQ64H", "MUFU.RCP64H", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "VOTE", "ATOMG", "CALL", "BSSY", "BRA", "UTCMMA", "LDTM", "UTMALDG"]
kern, counts, tot = None, {}, {}
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        tot[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        tot[kern] += 1
        op = m.group(1)
        for w in want:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
print("arch:", re.search(r"arch = (\S+)", sass).group(1))
for k in counts:
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name)
    print(f"\n{name}: {tot[k]} SASS instructions")
    print("  " + "  ".join(f"{w} {counts[k][w]}" for w in want if counts[k][w]))
