#!/bin/bash
# One gpurun call that validates HEAD on a B200: GPU tests, smoke, both bench arms.
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -n 4 gpurun_out/${tag}_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -n 2 gpurun_out/${tag}_smoke.log
( time timeout 400 python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench.json; tail -n 3 gpurun_out/${tag}_bench.err
( time timeout 300 python bench.py --impl reference ) > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 600 gpurun_out/${tag}_bench_ref.json
