#!/bin/bash
# full GPU test-suite with the bidiagonalisation SVD + phase cycles of the nT2 = 60 configurations
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for wl in cfg4gcv cfg4; do
  echo "[$wl]"; DECAES_PHASE_CYCLES=1 timeout 300 python bench.py --workload $wl --voxels 300000 --steps 1 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | grep -E "warp-cycles|^\{" | cut -c1-400
done
} 2>&1 | tee gpurun_out/r02_z3_pytest_phase.txt
