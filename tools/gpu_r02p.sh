#!/bin/bash
# L2 access-policy window: what the device grants, and DRAM traffic of a full-size launch for three policies
mkdir -p gpurun_out
DECAES_PHASE_CYCLES=1 timeout 100 python bench.py --voxels 100000 --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 2>&1 | grep "L2 window" | head -2
for e in "X=0" "DECAES_L2_MISS_NORMAL=1" "DECAES_NO_L2_WINDOW=1"; do
echo "== $e"
env $e ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:voxel_pipeline -c 1 --csv --log-file gpurun_out/r02p_traffic.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/r02p_traffic.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02p_traffic.csv')) if len(r)>10]
print('  '.join(f"{r[-3].split('.')[0]}={float(r[-1])/1e9:.2f}" for r in rows[1:]))
PY
done 2>&1 | tee gpurun_out/r02p_l2.txt
