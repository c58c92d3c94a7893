#!/bin/bash
# validation of HEAD (votes at the initial L-curve points): GPU tests, outputs with / without the votes, slab throughput
mkdir -p gpurun_out
{
( time timeout 700 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
for e in DECAES_STEP_SYNC=3 DECAES_NOP=1; do env $e python tools/out_hash.py 65536 lcurve 56 40 2>&1 | tail -1; env $e python tools/out_hash.py 8192 lcurve 48 60 2>&1 | tail -1; done
for r in 1 2; do echo -n "[cfg3] "; python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done
} 2>&1 | tee gpurun_out/r02_z23_validate_votes7.txt
