python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in 8 9; do for sm in 3 0; do echo "warps=$w sync=$sm"; DECAES_WARPS_PER_CTA=$w DECAES_SYNC_MASK=$sm DECAES_PHASE_CYCLES=1 python bench.py --voxels 400000 --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): print('  value', round(json.loads(l)['value']))
    elif 'warp-cycles' in l: print(' ', l.strip())
"; done; done
DECAES_LIB=$PWD/build/libdecaes_prof.so DECAES_PHASE_CYCLES=1 python bench.py --voxels 200000 --steps 1 --warmup 0 --no-e2e --no-cpu 2>&1 | grep -v "^{" | tail -19
