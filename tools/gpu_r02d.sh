#!/bin/bash
# batch 1: flip-angle probes without KKT polish / refinement; warps-per-SM scaling; identical-voxel volumes (divergence / i-cache test)
mkdir -p gpurun_out
run() { echo -n "[$1 $2] "; env $1 DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels ${VOX:-400000} --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 $2 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'chk', d['checksum_gdn'], t)
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; }
{
run "X=0"
run "DECAES_FA_POLISH=1"
run "DECAES_FA_REFINE=0"
run "X=0"
run "DECAES_FA_POLISH=1"
run "DECAES_FA_REFINE=0"
run "DECAES_WARPS_PER_CTA=4"
run "DECAES_WARPS_PER_CTA=8"
run "DECAES_WARPS_PER_CTA=10"
for k in 0 1 2 3 4 5; do run "X=0" "--identical $k"; done
for k in 0 1 2; do run "DECAES_WARPS_PER_CTA=4" "--identical $k"; done
} 2>&1 | tee gpurun_out/r02d_ab.txt
echo "--- parity, default"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_wide.py tests/test_golden.py -m gpu -q -x -s 2>&1 | grep -E "^(three|snr|one_pool|grid|nT2|gram vs|cfg1 full)|passed|failed|Error|error" | tee gpurun_out/r02d_parity_default.txt | tail -30
echo "--- parity, FA_REFINE=0"
DECAES_FA_REFINE=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_wide.py tests/test_golden.py -m gpu -q -s 2>&1 | grep -E "^(three|snr|one_pool|grid|nT2|gram vs|cfg1 full)|passed|failed|Error|error" | tee gpurun_out/r02d_parity_norefine.txt | tail -30
