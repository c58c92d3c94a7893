#!/bin/bash
# batch 11: one (column-major) copy of the voxel's basis in the global scratch, right-hand side from the EPG: A/B + full-size DRAM traffic
mkdir -p gpurun_out
run() { echo -n "[$1 $2] "; env $1 DECAES_PHASE_CYCLES=1 timeout 300 python bench.py --voxels ${VOX:-400000} --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 $2 2>&1 | python -c "
import sys,json
t='';p=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'chk', d['checksum_gdn'], t, p)
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; }
{
for r in 1 2; do
run "X=0"
run "DECAES_NEED_RM=1"
done
for wl in cfg1 cfg2 cfg4 cfg5; do run "X=0" "--workload $wl"; done
VOX=100000 run "X=0" "--workload cfg4gcv"
} 2>&1 | tee gpurun_out/r02o_ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_properties.py -m gpu -q 2>&1 | tail -3
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:voxel_pipeline -c 1 --csv --log-file gpurun_out/r02o_traffic_fullsize.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/r02o_traffic.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02o_traffic_fullsize.csv')) if len(r)>10]
for r in rows[1:]: print(r[-3], r[-2], r[-1])
PY
