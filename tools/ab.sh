# usage: ab.sh "ENV1=.. ENV2=.." "ENVA=.." ...   : alternates the given env settings, 3 rounds, 400k voxels
for r in 1 2 3; do for cfg in "$@"; do
  echo -n "[$cfg] "; env $cfg DECAES_PHASE_CYCLES=1 python bench.py --voxels ${VOX:-400000} --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
t=''
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'ms', round(d['ms_per_step'],1), 'kern_ms', round(d['kernel_ms_per_step'],1), t)
    elif 'warp-cycles' in l: t=l.strip().split('voxel:')[-1]
"; done; done
