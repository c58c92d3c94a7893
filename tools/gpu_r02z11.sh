#!/bin/bash
# gcv: trisection with four Sturm chains per lane (34 steps) against bisection with two (54 steps; libdecaes_prev.so = the commit before)
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
echo "== pytest gcv"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_wide.py -m gpu -q -k "gcv" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r02_z11_gcv_trisection.txt
