#!/bin/bash
# Profiling artefacts of one round (run under gpurun on ONE GPU): launch list of the bench command, a full
# ncu capture of the pipeline kernel on a 40k-voxel slab, and the DRAM traffic of one full-size launch.
tag=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --voxels 400000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voxel_pipeline -c 1 -o gpurun_out/${tag}_full \
    python bench.py --voxels 40000 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${tag}_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -k regex:voxel_pipeline -c 1 --csv --log-file gpurun_out/${tag}_traffic_fullsize.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${tag}_traffic.log 2>&1
tail -n 3 gpurun_out/${tag}_traffic_fullsize.csv
