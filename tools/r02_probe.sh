#!/bin/bash
# Round-2 kernel probe on one B200: parity tests, then the bank kernel in its variants.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest.txt
cat gpurun_out/r02_pytest.txt
{
for strat in slaney klapuri sampled; do
  echo "== $strat: tier on, tile group 2 (default)"; python tools/prof_bank.py $strat 4096 16384 6
  echo "== $strat: tier on, tile group 4"; ALZ_TILE_GROUP=4 python tools/prof_bank.py $strat 4096 16384 6
  echo "== $strat: tier off (all float64), group 2"; ALZ_NO_FP32_TIER=1 python tools/prof_bank.py $strat 4096 16384 6
done
echo "== slaney: all float32 forced (ALZ_TIER_TOL=1e9; NOT parity-safe, ceiling probe only), group 2 / 4"
ALZ_TIER_TOL=1e9 python tools/prof_bank.py slaney 4096 16384 6
ALZ_TIER_TOL=1e9 ALZ_TILE_GROUP=4 python tools/prof_bank.py slaney 4096 16384 6
echo "== slaney cfg5 shape 8192 x 8192"; python tools/prof_bank.py slaney 8192 8192 6
echo "== slaney real-time blocks 32768 x 480"; python tools/prof_bank.py slaney 32768 480 6
} 2>&1 | grep -v "^+" | tee gpurun_out/r02_probe.txt
