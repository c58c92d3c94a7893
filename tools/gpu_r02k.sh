#!/bin/bash
# batch 8: per-site votes (1 = flip-angle probes, 2 = L-curve steps, 4 = its initial points, 8 = Brent searches); full ncu capture
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
run "DECAES_STEP_SYNC=7"
run "DECAES_LC_HINTS=3"
run "DECAES_STEP_SYNC=2"
done
for wl in cfg1 cfg4 cfg5; do run "DECAES_STEP_SYNC=0" "--workload $wl"; run "DECAES_STEP_SYNC=1" "--workload $wl"; run "DECAES_STEP_SYNC=9" "--workload $wl";done
} 2>&1 | tee gpurun_out/r02k_ab.txt
ncu --set full --clock-control none --import-source on -k regex:voxel_pipeline -c 1 -o gpurun_out/r02k_full \
    python bench.py --voxels 100000 --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/r02k_full.log 2>&1
ls -la gpurun_out/r02k_full.ncu-rep
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02k_pytest.log 2>&1; tail -5 gpurun_out/r02k_pytest.log
