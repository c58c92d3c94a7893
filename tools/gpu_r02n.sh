#!/bin/bash
# validation of the round-2 kernel: GPU tests, smoke, both bench arms at full size, launch list, full ncu capture, full-size DRAM traffic
tag=${1:-r02n}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -s ) > gpurun_out/${tag}_pytest.log 2>&1
grep -E "^(three|snr|one_pool|grid|nT2|gram vs|cfg1 full)|passed|failed|FAILED" gpurun_out/${tag}_pytest.log | tail -40
timeout 120 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -n 1 gpurun_out/${tag}_smoke.log
( time timeout 500 python bench.py ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 700 gpurun_out/${tag}_bench.json; tail -n 3 gpurun_out/${tag}_bench.err
( time timeout 300 python bench.py --impl reference ) > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 300 gpurun_out/${tag}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voxel_pipeline -c 1 -o gpurun_out/${tag}_full \
    python bench.py --voxels 100000 --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/${tag}_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:voxel_pipeline -c 1 --csv --log-file gpurun_out/${tag}_traffic_fullsize.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/${tag}_traffic.log 2>&1
tail -n 3 gpurun_out/${tag}_traffic_fullsize.csv
