#!/bin/bash
# nT2 = 60 configurations: throughput against warps per SM (shared memory allows 6) - how much would more resident voxels buy?
mkdir -p gpurun_out
{
for wl in cfg5 cfg4; do for w in 3 4 5 6; do
  echo -n "[$wl warps=$w] "; DECAES_WARPS_PER_CTA=$w timeout 300 python bench.py --workload $wl --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done; done
echo "[cfg3 identical voxels]"; timeout 300 python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 --identical 7 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"
echo "[cfg3]"; timeout 300 python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"
} 2>&1 | tee gpurun_out/r02_z2_warps_nt2_60.txt
