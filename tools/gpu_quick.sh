#!/bin/bash
# quick GPU check: core parity tests + timing at 800k voxels (+ phase cycles)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_wide.py tests/test_golden.py -m gpu -q -x -s 2>&1 | grep -E "^(three|snr100|nT2|gram vs|cfg1 full)|passed|failed|Error|error" | head -30
DECAES_PHASE_CYCLES=1 timeout 200 python bench.py --voxels 800000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>&1 | grep -v "^{" | tail -n 3
for r in 1 2; do timeout 200 python bench.py --voxels 800000 --steps 2 --warmup 1 --no-e2e --no-cpu --parity-sample 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'kern_ms', round(d['kernel_ms_per_step'],1), 'chk', d['checksum_gdn'])
"; done
