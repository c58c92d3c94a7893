#!/bin/bash
# batch 10: vote groups / periods, polish variants, shared-memory Jacobi for Reg = gcv
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
run "DECAES_SYNC_GROUPS=2"
run "DECAES_SYNC_GROUPS=3"
run "DECAES_LC_VOTE_PERIOD=2"
run "DECAES_LC_VOTE_PERIOD=3"
run "DECAES_FA_POLISH=1"
run "DECAES_FA_POLISH=2"
run "DECAES_FA_POLISH=1 DECAES_KKT_TAU=1e-8"
done
VOX=100000 run "X=0" "--workload cfg4gcv"
VOX=100000 run "DECAES_GCV_SMEM=0" "--workload cfg4gcv"
} 2>&1 | tee gpurun_out/r02m_ab.txt
for e in "DECAES_FA_POLISH=2" "DECAES_FA_POLISH=1 DECAES_KKT_TAU=1e-8"; do
echo "--- parity $e"
env $e timeout 900 python -m pytest tests/test_gpu_parity_wide.py -m gpu -q -s -k "three or snr15 or nT2 or snr100" 2>&1 | grep -E "^(three|snr|one_pool|grid|nT2)|passed|failed|FAILED" | grep -v "chi2\|mdp" | tail -12
done 2>&1 | tee gpurun_out/r02m_parity_polish.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "gcv or golden" 2>&1 | tail -4
