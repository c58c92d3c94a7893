#!/bin/bash
# BASELINE configurations at their full sizes on one B200 (device-resident `value` + e2e through decaes_t2map); GPU tests of HEAD
mkdir -p gpurun_out
for wl in cfg1 cfg2 cfg4 cfg4gcv cfg5; do
  timeout 400 python bench.py --workload $wl --steps 2 --warmup 2 --no-cpu --parity-sample 2048 2> gpurun_out/r02_cfg_$wl.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print(json.dumps({'workload': d['config']['workload'], 'voxels': d['config']['voxels'], 'value': round(d['value']), 'unit': d['unit'], 'ms_per_step': round(d['ms_per_step'],1), 'e2e': round(d['e2e']['value']) if d.get('e2e') else None, 'e2e_pageable': round(d['e2e_pageable']['value']) if d.get('e2e_pageable') else None, 'parity': {k: d['parity'][k] for k in ('voxels_compared','out_of_tolerance_same_mu','mu_flips','mu_flips_between_two_cpu_builds','support_diff','early_returns','lcurve_overflow','nnls_itercap')} if d.get('parity') else None, 'clocks': d['clocks']}))
"
done 2>&1 | tee gpurun_out/r02_other_configs_fullsize.txt
( time timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -5
