#!/bin/bash
# batch 19: seed probes of the flip-angle fit without refinement / polish (re-probed precisely when they bracket the minimum)
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
for wl in cfg3 cfg3 cfg1 cfg2 cfg4 cfg5; do run "X=0" "--workload $wl"; run "DECAES_FA_ROUGH_SEEDS=0" "--workload $wl"; done
} 2>&1 | tee gpurun_out/r02x_ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_wide.py tests/test_golden.py -m gpu -q -s 2>&1 | grep -E "^(three|snr|one_pool|grid|nT2|gram vs|cfg1 full)|passed|failed|Error|error|FAILED" | tee gpurun_out/r02x_parity.txt | tail -34
