#!/bin/bash
# gram_factor (k > 8) with four entries per round, loads before stores (-DDECAES_FACTOR_CHUNKED) against the shipped loop
mkdir -p gpurun_out
A=$PWD/decaes.jl_b200/libdecaes_cuda.so
B=$PWD/decaes.jl_b200/libdecaes_chunk.so
{
for r in 1 2 3; do for L in $A $B; do
  echo -n "[$(basename $L)] "; DECAES_LIB=$L DECAES_PHASE_CYCLES=1 python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), t)
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; done; done
for L in $A $B; do DECAES_LIB=$L python tools/out_hash.py 65536 lcurve 56 40 2>&1 | tail -1; DECAES_LIB=$L python tools/out_hash.py 16384 lcurve 48 60 2>&1 | tail -1; done
for wl in cfg2 cfg5; do for L in $A $B; do
  echo -n "[$wl $(basename $L)] "; DECAES_LIB=$L python bench.py --workload $wl --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done; done
} 2>&1 | tee gpurun_out/r02_z21_factor_chunked.txt
