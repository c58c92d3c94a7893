#!/bin/bash
# knob sweep with the final kernel: phase barriers (DECAES_SYNC_MASK) and in-phase votes (DECAES_STEP_SYNC)
mkdir -p gpurun_out
run() { echo -n "[$1 $2] "; env $2 python bench.py --workload $1 --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; }
{
for e in DECAES_NOP=1 DECAES_SYNC_MASK=3 DECAES_SYNC_MASK=5 DECAES_STEP_SYNC=7 DECAES_FA_WARM=8 DECAES_NOP=2; do run cfg3 $e; done
for e in DECAES_NOP=1 DECAES_SYNC_MASK=3 DECAES_SYNC_MASK=1 DECAES_FA_WARM=8; do run cfg5 $e; done
for e in DECAES_NOP=1 DECAES_SYNC_MASK=3 DECAES_SYNC_MASK=1; do run cfg4 $e; done
} 2>&1 | tee gpurun_out/r02_z20_knobs.txt
