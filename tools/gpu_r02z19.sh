#!/bin/bash
# validation of HEAD (votes before the flip-angle probes at six warps per SM): GPU tests + nT2 = 60 throughput + determinism of outputs with / without votes
mkdir -p gpurun_out
{
( time timeout 700 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
for e in DECAES_STEP_SYNC=0 DECAES_NOP=1; do env $e python tools/out_hash.py 16384 mdp 32 60 2>&1 | tail -1; env $e python tools/out_hash.py 16384 chi2 48 60 2>&1 | tail -1; env $e python tools/out_hash.py 8192 lcurve 48 60 2>&1 | tail -1; done
for wl in cfg4 cfg5; do echo -n "[$wl] "; python bench.py --workload $wl --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 2048 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); p=d.get('parity') or {}; print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'flips', p.get('mu_flips'), 'same_mu_out', p.get('out_of_tolerance_same_mu'))
"; done
} 2>&1 | tee gpurun_out/r02_z19_validate_votes6.txt
