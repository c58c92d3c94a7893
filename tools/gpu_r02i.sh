#!/bin/bash
# batch 6: direct Cholesky solve for the full-set start; CTA votes inside the phases (per L-curve step / per flip-angle probe)
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
run "DECAES_LC_HINTS=3"
run "DECAES_LC_SYNC=1"
run "DECAES_FA_SYNC=1"
run "DECAES_LC_SYNC=1 DECAES_FA_SYNC=1"
done
} 2>&1 | tee gpurun_out/r02i_ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_wide.py tests/test_golden.py -m gpu -q -s 2>&1 | grep -E "^(three|snr|one_pool|grid|nT2|gram vs|cfg1 full)|passed|failed|Error|error|FAILED" | tee gpurun_out/r02i_parity.txt | tail -32
