for w in 4 6 8; do for sm in 3 0; do echo "warps=$w sync=$sm"; DECAES_WARPS_PER_CTA=$w DECAES_SYNC_MASK=$sm DECAES_PHASE_CYCLES=1 python bench.py --voxels 400000 --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): print('  value', round(json.loads(l)['value']))
    elif 'warp-cycles' in l: print(' ', l.strip())
"; done; done
