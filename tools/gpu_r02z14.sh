#!/bin/bash
# echo trains of 80 and 96 echoes (nTE limit raised from 72 to 96) against the oracle; headline unchanged (hash + slab throughput)
mkdir -p gpurun_out
{
timeout 800 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "odd_sizes" 2>&1 | tail -12
python tools/out_hash.py 65536 lcurve 56 40 2>&1 | tail -1
for r in 1 2; do echo -n "[cfg3] "; python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done
} 2>&1 | tee gpurun_out/r02_z14_nte96.txt
