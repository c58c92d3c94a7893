#!/bin/bash
# round 2, GPU call A: sanity tests, L-curve parity A/B (65,536 voxels), phase/function cycle profile, refine cost, TMEM latency
tag=r02a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nproc > gpurun_out/${tag}_nproc.txt
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; tail -n 3 gpurun_out/${tag}_pytest.log
timeout 60 build/tmem_lat > gpurun_out/${tag}_tmem_lat.txt 2>&1; cat gpurun_out/${tag}_tmem_lat.txt
timeout 900 python tools/lcurve_ab.py --voxels 65536 --out gpurun_out/${tag}_lcurve_ab.json > gpurun_out/${tag}_lcurve_ab.log 2>&1; tail -n 3 gpurun_out/${tag}_lcurve_ab.log
for R in 0 1; do
  DECAES_REFINE=$R DECAES_LIB=build/libdecaes_prof.so DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels 200000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${tag}_prof_refine$R.log 2>&1
done
grep -A3 "warp-cycles" gpurun_out/${tag}_prof_refine0.log | tail -n 8
for r in 1 2; do for R in 0 1; do
  echo -n "[REFINE=$R] "; DECAES_REFINE=$R timeout 120 python bench.py --voxels 800000 --steps 2 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'chk', d['checksum_gdn'])
"; done; done 2>&1 | tee gpurun_out/${tag}_ab_refine.txt
