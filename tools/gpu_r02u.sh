#!/bin/bash
# batch 16: dilated warm starts for the second / third L-curve point; EPG state loop unrolled twice; masked full-size volume
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
run "DECAES_LC_HINTS=15"
run "DECAES_LIB=build/libdecaes_eu2.so"
done
run "DECAES_LC_HINTS=15" "--workload cfg2"
run "X=0" "--workload cfg2"
} 2>&1 | tee gpurun_out/r02u_ab.txt
DECAES_LC_HINTS=15 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "lcurve or golden or oracle" 2>&1 | tail -3
timeout 300 python bench.py --mask 0.65 --steps 2 --warmup 2 --no-cpu --parity-sample 0 > gpurun_out/r02u_bench_masked.json 2> gpurun_out/r02u_bench_masked.err; tail -c 600 gpurun_out/r02u_bench_masked.json | head -c 600
