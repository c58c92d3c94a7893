#!/bin/bash
# votes at the four initial L-curve points too (DECAES_STEP_SYNC=7) against the default (3), three rounds
mkdir -p gpurun_out
run() { echo -n "[$1 $2] "; env $2 python bench.py --workload $1 --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; }
{
for r in 1 2 3; do for e in DECAES_STEP_SYNC=3 DECAES_STEP_SYNC=7; do run cfg3 $e; done; done
for e in DECAES_STEP_SYNC=3 DECAES_STEP_SYNC=7; do run cfg2 $e; done
} 2>&1 | tee gpurun_out/r02_z22_votes_initial_points.txt
