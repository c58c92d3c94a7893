#!/bin/bash
# validation of HEAD (warm starts from ones): GPU tests, smoke, bench at full size
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_s5_pytest.log 2>&1; tail -n 5 gpurun_out/r02_s5_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02_s5_smoke.log 2>&1; tail -n 1 gpurun_out/r02_s5_smoke.log | cut -c1-200
timeout 400 python bench.py > gpurun_out/r02_s5_bench.json 2> gpurun_out/r02_s5_bench.err; tail -c 300 gpurun_out/r02_s5_bench.json
