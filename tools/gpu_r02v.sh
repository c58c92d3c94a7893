#!/bin/bash
# batch 17: gram_build with four L2 round trips per tile row (unroll 4), EPG unroll 2 (default) vs 4; prev = last commit
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
for r in 1 2; do
run "X=0"
run "DECAES_LIB=build/libdecaes_eu4.so"
run "DECAES_LIB=build/libdecaes_prev.so"
done
for wl in cfg1 cfg2 cfg4 cfg5; do run "X=0" "--workload $wl"; run "DECAES_LIB=build/libdecaes_prev.so" "--workload $wl"; done
} 2>&1 | tee gpurun_out/r02v_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q 2>&1 | tail -3
