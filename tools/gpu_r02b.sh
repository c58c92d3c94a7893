#!/bin/bash
tag=r02b
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x --durations=12 -s ) > gpurun_out/${tag}_pytest.log 2>&1; tail -n 25 gpurun_out/${tag}_pytest.log
DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels 800000 --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | grep -v "^{" | tail -n 3
