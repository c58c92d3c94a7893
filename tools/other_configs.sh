#!/bin/bash
# throughput of the non-headline configurations on 500k-voxel slabs (debug override of bench.py, not a headline number)
mkdir -p gpurun_out
for wl in cfg1 cfg2 cfg4 cfg5; do
  echo -n "[$wl] "; timeout 60 python bench.py --workload $wl --voxels 500000 --steps 2 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done 2>&1 | tee gpurun_out/other_configs.txt
