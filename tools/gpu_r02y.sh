#!/bin/bash
for e in "DECAES_FA_ROUGH_SEEDS=1" "DECAES_FA_ROUGH_SEEDS=0"; do
echo "== $e"
env $e timeout 600 python -m pytest tests/test_gpu_parity_wide.py -m gpu -q -s -k "snr15 or snr25" 2>&1 | grep -E "^(snr)|passed|failed|AssertionError" | cut -c1-900
done
