#!/bin/bash
# full ncu capture of the pipeline kernel on a 100k-voxel slab (one launch); env LIBTAG picks a build/ variant
tag=${1:-r02x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:voxel_pipeline -c 1 -o gpurun_out/${tag}_full \
    python bench.py --voxels 100000 --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/${tag}_full.log 2>&1
ls -la gpurun_out/${tag}_full.ncu-rep
