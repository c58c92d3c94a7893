#!/bin/bash
# seed probes of the flip-angle fit warm-started from the previous voxel's active set at the same angle (DECAES_FA_PREV=1)
mkdir -p gpurun_out
{
for r in 1 2; do for e in DECAES_FA_PREV=0 DECAES_FA_PREV=1; do
  echo -n "[$e] "; env $e DECAES_PHASE_CYCLES=1 python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample $([ $r = 1 ] && echo 8192 || echo 0) 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); p=d.get('parity') or {}; print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), t, 'flips', p.get('mu_flips'), 'same_mu_out', p.get('out_of_tolerance_same_mu'), 'cpu-cpu', p.get('mu_flips_between_two_cpu_builds'), 'alpha', p.get('alpha_max_abs'))
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; done; done
for wl in cfg1 cfg5; do for e in DECAES_FA_PREV=0 DECAES_FA_PREV=1; do
  echo -n "[$wl $e] "; env $e python bench.py --workload $wl --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 2048 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); p=d.get('parity') or {}; print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'flips', p.get('mu_flips'), 'same_mu_out', p.get('out_of_tolerance_same_mu'))
"; done; done
} 2>&1 | tee gpurun_out/r02_z17_fa_prev.txt
