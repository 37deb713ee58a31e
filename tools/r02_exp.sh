#!/bin/bash
# Where is the 3.45 ms ceiling?  ALZ_EXP: 1 = load only the first tile group (no L2 read stream), 2 = no tile stores.
mkdir -p gpurun_out
{
for tol in 1e9 2.5e-6; do
  for exp in 0 1 2 3; do
    echo "== slaney tol=$tol ALZ_EXP=$exp (group 2)"; ALZ_TIER_TOL=$tol ALZ_EXP=$exp python tools/prof_bank.py slaney 4096 16384 5
  done
done
echo "== slaney tol=1e9 ALZ_EXP=1 group 4"; ALZ_TIER_TOL=1e9 ALZ_EXP=1 ALZ_TILE_GROUP=4 python tools/prof_bank.py slaney 4096 16384 5
echo "== slaney tol=1e9 ALZ_EXP=0 group 1 (prefetch pipeline)"; ALZ_TIER_TOL=1e9 ALZ_TMA_PAIRED=1 python tools/prof_bank.py slaney 4096 16384 5
} 2>&1 | tee gpurun_out/r02_exp.txt
