#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "legacy" -s ) > gpurun_out/c2_legacy.log 2>&1
tail -n 30 gpurun_out/c2_legacy.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c2_pytest.log 2>&1
tail -n 5 gpurun_out/c2_pytest.log
for r in 1 2; do for lib in build/libdecaes_base.so decaes.jl_b200/libdecaes_cuda.so; do
  echo -n "[$lib] "; DECAES_LIB=$PWD/$lib timeout 120 python bench.py --voxels 800000 --steps 2 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done; done 2>&1 | tee gpurun_out/c2_ab.txt
