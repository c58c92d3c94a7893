#!/bin/bash
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q -k "pathological or release" ) > gpurun_out/c4_pytest.log 2>&1
tail -n 6 gpurun_out/c4_pytest.log
run() { echo -n "[$1 $2] "; env $2 DECAES_LIB=$PWD/$1 timeout 120 python bench.py --voxels 800000 --steps 2 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'chk', d['checksum_gdn'])
"; }
for r in 1 2; do
  run decaes.jl_b200/libdecaes_cuda.so X=1
  run build/libdecaes_u2.so X=1
  run build/libdecaes_u4.so X=1
  run decaes.jl_b200/libdecaes_cuda.so DECAES_SYNC_MASK=5
  run decaes.jl_b200/libdecaes_cuda.so DECAES_SYNC_MASK=6
done 2>&1 | tee gpurun_out/c4_ab.txt
