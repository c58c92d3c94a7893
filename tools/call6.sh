#!/bin/bash
mkdir -p gpurun_out
run() { echo -n "[$1] "; env $1 timeout 120 python bench.py --voxels 800000 --steps 2 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'chk', d['checksum_gdn'])
"; }
for r in 1 2; do
  run DECAES_GV_STRIDE_64=1
  run X=1
done 2>&1 | tee gpurun_out/c6_ab.txt
( time timeout 500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c6_pytest.log 2>&1
tail -n 6 gpurun_out/c6_pytest.log
