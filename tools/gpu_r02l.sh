#!/bin/bash
# batch 9: L-curve curvature bookkeeping through sorted-order links (must be bit-identical: same checksum); sync policy by CTA shape; ncu
mkdir -p gpurun_out
run() { echo -n "[$1 $2] "; env $1 DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels ${VOX:-400000} --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 $2 2>&1 | python -c "
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
done
for wl in cfg1 cfg2 cfg4 cfg5; do run "X=0" "--workload $wl"; done
} 2>&1 | tee gpurun_out/r02l_ab.txt
ncu --set full --clock-control none --import-source on -k regex:voxel_pipeline -c 1 -o gpurun_out/r02l_full \
    python bench.py --voxels 100000 --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/r02l_full.log 2>&1
ls -la gpurun_out/r02l_full.ncu-rep
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q 2>&1 | tail -3
