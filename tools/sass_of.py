#!/usr/bin/env python
"""Print the SASS of one out-of-line device function of a kernel (by substring of its demangled name)."""
import re, subprocess, sys
so = "decaes.jl_b200/libdecaes_cuda.so"
kern = sys.argv[2] if len(sys.argv) > 2 else "voxel_pipeline_kernelILb1ELb0ELi40E"
want = sys.argv[1]
elf = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
syms = []
for line in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)", line)
    if m and kern in m.group(3):
        syms.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
names = subprocess.run(["c++filt"] + [s[2] for s in syms], capture_output=True, text=True).stdout.splitlines()
sel = [(o, sz, n) for (o, sz, _), n in zip(syms, names) if want in n]
if not sel:
    sys.exit("no such function; have: " + ", ".join(sorted(set(re.sub(r"\(.*", "", n) for n in names))))
off, size, name = sel[0]
print("//", name, "offset", hex(off), "size", size // 16, "instr")
sass = subprocess.run(["cuobjdump", "-sass", "-fun", [l for l in subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout.split() if kern in l and l.startswith(".text.")][0][6:], so], capture_output=True, text=True).stdout
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*/\*", line)
    if m and off <= int(m.group(1), 16) < off + size:
        print(f"{int(m.group(1),16)-off:5x}  {m.group(2)}")
