#!/bin/bash
# batch 20: fa_probe out of line, rough-seed code compiled out (optimize_flip_angle 824 -> 180 instructions); prev = last commit
mkdir -p gpurun_out
run() { echo -n "[$1 $2] "; env $1 DECAES_PHASE_CYCLES=1 timeout 300 python bench.py --voxels ${VOX:-400000} --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 $2 2>&1 | python -c "
import sys,json
t='';p=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'chk', d['checksum_gdn'], t, p)
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; }
{
for wl in cfg3 cfg3 cfg1 cfg4 cfg5; do run "X=0" "--workload $wl"; run "DECAES_LIB=build/libdecaes_r02z.so" "--workload $wl"; done
} 2>&1 | tee gpurun_out/r02aa_ab.txt
