#!/usr/bin/env python
"""SASS instruction count of every out-of-line device function of voxel_pipeline_kernel<true>."""
import re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "decaes.jl_b200/libdecaes_cuda.so"
kern = sys.argv[2] if len(sys.argv) > 2 else "voxel_pipeline_kernelILb1ELb0E"
elf = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
rows = []
for line in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", line)
    if m and kern in m.group(3):
        rows.append((int(m.group(2), 16) // 16, m.group(4)))
    m = re.match(r"\s*[0-9a-f]+\s+[0-9a-f]+\s+([0-9a-f]+)\s.*PROGBITS.*\.text\.(\S+)", line)
    if m and kern in m.group(2):
        total = int(m.group(1), 16) // 16
rows = sorted(set(rows), reverse=True)
names = subprocess.run(["c++filt"] + [r[1] for r in rows], capture_output=True, text=True).stdout.splitlines()
print("kernel text total:", total, "instr =", total * 16 // 1024, "KB; out-of-line functions:", sum(r[0] for r in rows))
for (n, _), nm in zip(rows, names):
    print(f"{n:6d}  {re.sub(r'\(.*', '', nm).replace('decaes::', '')}")
