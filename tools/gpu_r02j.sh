#!/bin/bash
# batch 7: CTA votes at every NNLS solve (DECAES_STEP_SYNC=1, default) / also at every active-set iteration (=3); re-tests under sync
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
run "DECAES_STEP_SYNC=0"
run "DECAES_STEP_SYNC=1"
run "DECAES_STEP_SYNC=3"
run "DECAES_LIB=build/libdecaes_fs.so"
run "DECAES_LC_HINTS=7"
run "DECAES_FA_POLISH=1"
done
for wl in cfg1 cfg2 cfg4 cfg5; do run "DECAES_STEP_SYNC=0" "--workload $wl"; run "DECAES_STEP_SYNC=1" "--workload $wl"; done
} 2>&1 | tee gpurun_out/r02j_ab.txt
M=gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:voxel_pipeline -c 1 --csv --log-file gpurun_out/r02j_icache.csv \
    python bench.py --voxels 100000 --steps 1 --warmup 0 --no-e2e --no-cpu --parity-sample 0 > gpurun_out/r02j_icache.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02j_icache.csv')) if len(r)>10]
for r in rows[1:]: print(r[-3], r[-1])
PY
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02j_pytest.log 2>&1; tail -5 gpurun_out/r02j_pytest.log
