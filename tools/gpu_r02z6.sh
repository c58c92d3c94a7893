#!/bin/bash
# register-resident factorisation for 8 < k <= 16 (gram_factor_mid) against the shared-memory one (-DDECAES_NO_FACTOR_MID build)
mkdir -p gpurun_out
A=$PWD/decaes.jl_b200/libdecaes_nomid.so
B=$PWD/decaes.jl_b200/libdecaes_cuda.so
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
echo "== bit equality of the outputs"
for L in $A $B; do DECAES_LIB=$L python tools/out_hash.py 65536 lcurve 56 40; DECAES_LIB=$L python tools/out_hash.py 16384 chi2 48 60; DECAES_LIB=$L python tools/out_hash.py 16384 gcv 48 60; done
echo "== other configs"
for wl in cfg2 cfg4 cfg5; do for L in $A $B; do
  echo -n "[$wl $(basename $L)] "; DECAES_LIB=$L python bench.py --workload $wl --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done; done
} 2>&1 | tee gpurun_out/r02_z6_factor_mid.txt
