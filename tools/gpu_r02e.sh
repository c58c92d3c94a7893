#!/bin/bash
# batch 2: compact kkt_pick (polish on/off on the probes), L-curve start hints, parity
mkdir -p gpurun_out
run() { echo -n "[$1 $2] "; env $1 DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels ${VOX:-400000} --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 $2 2>&1 | python -c "
import sys,json
t='';p=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'chk', d['checksum_gdn'], t, p)
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
    elif 'KKT-polish' in l: p='polish/voxel '+l.strip().split(':')[-1]
"; }
{
for r in 1 2; do
run "DECAES_LC_HINTS=0"
run "DECAES_LC_HINTS=1"
run "DECAES_LC_HINTS=2"
run "DECAES_LC_HINTS=3"
run "DECAES_FA_POLISH=1"
done
for wl in cfg1 cfg2 cfg4 cfg4gcv cfg5; do run "X=0" "--workload $wl"; done
} 2>&1 | tee gpurun_out/r02e_ab.txt
echo "--- parity, default"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_wide.py tests/test_golden.py tests/test_gpu_properties.py -m gpu -q -s 2>&1 | grep -E "^(three|snr|one_pool|grid|nT2|gram vs|cfg1 full)|passed|failed|Error|error|FAILED" | tee gpurun_out/r02e_parity_default.txt | tail -40
echo "--- parity, FA_POLISH=1"
DECAES_FA_POLISH=1 timeout 900 python -m pytest tests/test_gpu_parity_wide.py -m gpu -q -s -k "three or snr15 or nT2" 2>&1 | grep -E "^(three|snr|one_pool|grid|nT2|gram vs|cfg1 full)|passed|failed|Error|error|FAILED" | tee gpurun_out/r02e_parity_polish.txt | tail -30
