#!/bin/bash
# batch 3: register-resident small factorisation, merged append pass, EPG pass split 24 + 16; profile build; full GPU tests
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
run "DECAES_LIB=build/libdecaes_nofs.so"
run "DECAES_LIB=build/libdecaes_oldapp.so"
run "DECAES_EPG_LANES=20"
done
for wl in cfg1 cfg2 cfg4 cfg5; do run "X=0" "--workload $wl"; done
} 2>&1 | tee gpurun_out/r02f_ab.txt
for e in "X=0" "DECAES_FA_POLISH=1"; do
echo "--- profile build $e"
env $e DECAES_LIB=build/libdecaes_prof.so DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels 200000 --steps 1 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | grep -v "^{" | head -70
done > gpurun_out/r02f_profile.txt 2>&1
echo "--- GPU tests"
( time timeout 900 python -m pytest tests -m gpu -q -s ) > gpurun_out/r02f_pytest.log 2>&1
grep -E "^(three|snr|one_pool|grid|nT2|gram vs|cfg1 full)|passed|failed|Error|error|FAILED" gpurun_out/r02f_pytest.log | tail -45
