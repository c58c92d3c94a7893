#!/bin/bash
# gcv: two Sturm pivots per reciprocal + branch-free reciprocal (DECAES_GCV_STURM2=1) against one pivot per __drcp_rn (=0);
# fresh per-solve histogram of the headline config (DECAES_PROFILE build)
mkdir -p gpurun_out
{
for r in 1 2; do for e in DECAES_GCV_STURM2=0 DECAES_GCV_STURM2=1; do
  echo -n "[$e] "; env $e DECAES_PHASE_CYCLES=1 timeout 300 python bench.py --workload cfg4gcv --voxels 300000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 2048 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); p=d.get('parity') or {}; print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), t, 'flips', p.get('mu_flips'), 'same_mu_out', p.get('out_of_tolerance_same_mu'), 'cpu-cpu flips', p.get('mu_flips_between_two_cpu_builds'))
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; done; done
echo "== pytest gcv"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gcv" 2>&1 | tail -3
echo "== profile build, cfg3, 200k voxels"
DECAES_LIB=$PWD/decaes.jl_b200/libdecaes_prof.so DECAES_PHASE_CYCLES=1 timeout 300 python bench.py --voxels 200000 --steps 1 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | grep -v "^{" | tail -75
} 2>&1 | tee gpurun_out/r02_z4_gcv_sturm2_profile.txt
