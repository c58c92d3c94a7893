#!/bin/bash
# shared-memory EPG: last pass (16 components) on two lanes per component (DECAES_EPG_SPLIT=1, default) against one lane (=0)
mkdir -p gpurun_out
{
for r in 1 2 3; do for e in DECAES_EPG_SPLIT=0 DECAES_EPG_SPLIT=1; do
  echo -n "[$e] "; env $e DECAES_PHASE_CYCLES=1 python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), t)
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; done; done
echo "== bit equality of the outputs"
for e in DECAES_EPG_SPLIT=0 DECAES_EPG_SPLIT=1; do env $e python tools/out_hash.py 65536 lcurve 56 40 2>&1 | tail -1; env $e python tools/out_hash.py 16384 none 32 40 2>&1 | tail -1; env $e python tools/out_hash.py 16384 chi2 48 40 2>&1 | tail -1; done
echo "== cfg1 / cfg2"
for wl in cfg1 cfg2; do for e in DECAES_EPG_SPLIT=0 DECAES_EPG_SPLIT=1; do
  echo -n "[$wl $e] "; env $e python bench.py --workload $wl --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done; done
} 2>&1 | tee gpurun_out/r02_z12_epg_split.txt
