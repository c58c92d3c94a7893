#!/bin/bash
# batch 4: EPG passes with idle lanes sitting out (24 + 16 vs 20 + 20), full ncu capture of the pipeline kernel (per-function stalls)
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
run "DECAES_EPG_LANES=20"
run "DECAES_EPG_LANES=32"
done
} 2>&1 | tee gpurun_out/r02g_ab.txt
ncu --set full --clock-control none --import-source on -k regex:voxel_pipeline -c 1 -o gpurun_out/r02g_full \
    python bench.py --voxels 100000 --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/r02g_full.log 2>&1
ls -la gpurun_out/r02g_full.ncu-rep
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
