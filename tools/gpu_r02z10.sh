#!/bin/bash
# gcv: reflector application through reflect_apply (no second-vector loads when no lane has one) against the previous commit (libdecaes_prev.so)
mkdir -p gpurun_out
A=$PWD/decaes.jl_b200/libdecaes_prev.so
B=$PWD/decaes.jl_b200/libdecaes_cuda.so
{
for r in 1 2; do for L in $A $B; do
  echo -n "[cfg4gcv $(basename $L)] "; DECAES_LIB=$L DECAES_PHASE_CYCLES=1 timeout 300 python bench.py --workload cfg4gcv --voxels 300000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 2048 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); p=d.get('parity') or {}; print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), t, 'flips', p.get('mu_flips'), 'same_mu_out', p.get('out_of_tolerance_same_mu'), 'cpu-cpu flips', p.get('mu_flips_between_two_cpu_builds'))
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; done; done
for L in $A $B; do DECAES_LIB=$L python tools/out_hash.py 16384 gcv 48 60 2>&1 | tail -1; DECAES_LIB=$L python tools/out_hash.py 16384 gcv 56 40 2>&1 | tail -1;  DECAES_LIB=$L python tools/out_hash.py 65536 lcurve 56 40 2>&1 | tail -1; done
echo "== pytest"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
} 2>&1 | tee gpurun_out/r02_z10_gcv_reflect_apply.txt
