#!/bin/bash
# Store-stream experiments: st.global instead of TMA stores (ALZ_EXP=4); channel-major output layout y[C][S][T]
mkdir -p gpurun_out
{
for v in "ALZ_TIER_TOL=1e9" "ALZ_TIER_TOL=1e9 ALZ_EXP=4" "ALZ_TIER_TOL=1e9 ALZ_PROF_LAYOUT=channel" "ALZ_X=0" "ALZ_EXP=4" "ALZ_PROF_LAYOUT=channel" "ALZ_PROF_LAYOUT=channel ALZ_NO_FP32_TIER=1"; do
  echo "== $v"; env $v timeout 60 python tools/prof_bank.py slaney 4096 16384 5
done
echo "== klapuri channel-major"; ALZ_PROF_LAYOUT=channel timeout 60 python tools/prof_bank.py klapuri 4096 16384 5
echo "== sampled channel-major"; ALZ_PROF_LAYOUT=channel timeout 60 python tools/prof_bank.py sampled 4096 16384 5
} 2>&1 | tee gpurun_out/r02_exp3.txt
