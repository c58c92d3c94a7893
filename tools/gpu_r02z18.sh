#!/bin/bash
# nT2 = 60 configurations (six warps per SM): votes before the flip-angle probes on / off with the final kernel
mkdir -p gpurun_out
{
for wl in cfg4 cfg5 cfg4gcv; do for e in DECAES_STEP_SYNC=0 DECAES_STEP_SYNC=1; do
  echo -n "[$wl $e] "; env $e python bench.py --workload $wl --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done; done
} 2>&1 | tee gpurun_out/r02_z18_votes_nt2_60.txt
