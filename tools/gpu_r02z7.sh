#!/bin/bash
# L-curve curvature update with the four state points in parallel (default build) against the sequential version (-DDECAES_LC_CURV_SEQ),
# and eight columns per round trip in the explicit residual for nT2 <= 40 only (-DDECAES_RESID_CHUNK_VS40=8)
mkdir -p gpurun_out
A=$PWD/decaes.jl_b200/libdecaes_seq.so
B=$PWD/decaes.jl_b200/libdecaes_cuda.so
Cc=$PWD/decaes.jl_b200/libdecaes_c8.so
{
for r in 1 2 3; do for L in $A $B $Cc; do
  echo -n "[$(basename $L)] "; DECAES_LIB=$L DECAES_PHASE_CYCLES=1 python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), t)
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; done; done
echo "== bit equality of the outputs"
for L in $A $B $Cc; do DECAES_LIB=$L python tools/out_hash.py 65536 lcurve 56 40 2>&1 | tail -1; DECAES_LIB=$L python tools/out_hash.py 16384 lcurve 48 60 2>&1 | tail -1; done
echo "== cfg2"
for L in $A $B $Cc; do
  echo -n "[cfg2 $(basename $L)] "; DECAES_LIB=$L python bench.py --workload cfg2 --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done
} 2>&1 | tee gpurun_out/r02_z7_curv_par_chunk8.txt
