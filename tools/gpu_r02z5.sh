#!/bin/bash
# gcv: unnormalised reflectors, chunked loads in the bidiagonalisation, pipelined copy (against the commit before: 553 k voxels/s)
mkdir -p gpurun_out
{
for r in 1 2; do
  echo -n "[cfg4gcv] "; DECAES_PHASE_CYCLES=1 timeout 300 python bench.py --workload cfg4gcv --voxels 300000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 2048 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); p=d.get('parity') or {}; print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), t, 'flips', p.get('mu_flips'), 'same_mu_out', p.get('out_of_tolerance_same_mu'), 'cpu-cpu flips', p.get('mu_flips_between_two_cpu_builds'))
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; done
echo "== pytest gcv + fixed angle"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gcv" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r02_z5_gcv_bidiag2.txt
