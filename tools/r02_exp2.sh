#!/bin/bash
# Store-stream ceiling vs row stride (all channels forced to float32: arithmetic ~2 ms, so the memory side shows).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest.txt; cat gpurun_out/r02_pytest.txt
{
for T in 16384 16352 16416 12288 20480 8192 32768; do
  S=$((4096*16384/T/32*32))
  echo "== all-f32, T=$T S=$S"; ALZ_TIER_TOL=1e9 python tools/prof_bank.py slaney $S $T 5
done
echo "== all-f32, cp.async engine (st.global.v4 stores)"; ALZ_TIER_TOL=1e9 ALZ_NO_TMA=1 python tools/prof_bank.py slaney 4096 16384 5
echo "== tiered, tol 3.3e-6"; ALZ_TIER_TOL=3.3e-6 python tools/prof_bank.py slaney 4096 16384 5
echo "== tiered default"; python tools/prof_bank.py slaney 4096 16384 5
echo "== tiered default, 32 channels"; python tools/prof_bank.py slaney 8192 16384 5 32
echo "== tiered default, no segments"; ALZ_NO_SEGMENT=1 python tools/prof_bank.py slaney 4096 16384 5
echo "== tiered default, seg waves 8 / 32"; ALZ_SEG_WAVES=8 python tools/prof_bank.py slaney 4096 16384 5; ALZ_SEG_WAVES=32 python tools/prof_bank.py slaney 4096 16384 5
} 2>&1 | tee gpurun_out/r02_exp2.txt
