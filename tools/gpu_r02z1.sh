#!/bin/bash
# Reg = gcv: singular values by bidiagonalisation + bisection (DECAES_GCV_SMEM=2, new default) against the parallel Jacobi (=1)
mkdir -p gpurun_out
{
for r in 1 2; do for e in DECAES_GCV_SMEM=1 DECAES_GCV_SMEM=2; do
  echo -n "[$e] "; env $e timeout 300 python bench.py --workload cfg4gcv --voxels 300000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 2048 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'parity', json.dumps(d.get('parity')))
"; done; done
echo "== pytest gcv"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gcv" 2>&1 | tail -5
} 2>&1 | tee gpurun_out/r02_z1_gcv_svd.txt
