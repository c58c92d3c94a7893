#!/bin/bash
# batch 5: instruction-cache diet (no auto-unrolled mask_of, save_results loops not unrolled, smaller residual / atv), icache metrics
mkdir -p gpurun_out
run() { echo -n "[$1 $2] "; env $1 DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels ${VOX:-400000} --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 $2 2>&1 | python -c "
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
run "DECAES_FA_WARM=0"
done
} 2>&1 | tee gpurun_out/r02h_ab.txt
M=gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:voxel_pipeline -c 1 --csv --log-file gpurun_out/r02h_icache.csv \
    python bench.py --voxels 100000 --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/r02h_icache.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02h_icache.csv')) if len(r)>10]
for r in rows[1:]: print(r[-3], r[-1])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
