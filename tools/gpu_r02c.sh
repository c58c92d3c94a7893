#!/bin/bash
tag=r02c
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --durations=8 -s ) > gpurun_out/${tag}_pytest.log 2>&1; tail -n 16 gpurun_out/${tag}_pytest.log
grep -E "^(snr|one_pool|three|grid|nT2|gram vs|cfg1 full)" gpurun_out/${tag}_pytest.log
DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels 800000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | grep -v "^{" | tail -n 3
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 3000 gpurun_out/${tag}_bench.json; tail -n 5 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 600 gpurun_out/${tag}_bench_ref.json
