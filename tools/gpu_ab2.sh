#!/bin/bash
run() { echo -n "[$1] "; env $1 DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels ${VOX:-400000} --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), t)
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; }
run "X=0"
run "DECAES_WARPS_PER_CTA=12"
run "DECAES_WARPS_PER_CTA=8"
run "DECAES_LIB=build/libdecaes_w12.so"
run "DECAES_LIB=build/libdecaes_w12.so DECAES_WARPS_PER_CTA=8"
for L in prof prof12; do
DECAES_LIB=build/libdecaes_$L.so DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels 200000 --steps 1 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | grep -v "^{" | head -24
done
