#!/bin/bash
# One gpurun call that validates HEAD on a B200: GPU tests, smoke, both bench arms, a warps-per-SM A/B
# (DECAES_SPILL = 0/1/3 -> 9/10/11 warps on cfg3) and the profiling artefacts of tools/profile_round.sh.
tag=${1:-r01s3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
( time timeout 540 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -n 4 gpurun_out/${tag}_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -n 2 gpurun_out/${tag}_smoke.log
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json
timeout 200 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 400 gpurun_out/${tag}_bench_ref.json
for r in 1; do for sp in 0 1 3; do
  echo -n "[SPILL=$sp] "; DECAES_SPILL=$sp timeout 120 python bench.py --voxels 800000 --steps 2 --warmup 1 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1))
"; done; done 2>&1 | tee gpurun_out/${tag}_ab_spill.txt
timeout 420 bash tools/profile_round.sh ${tag}
