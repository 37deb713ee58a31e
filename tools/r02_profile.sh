#!/bin/bash
# Round-2 evidence on one B200: tests, bench line, ncu launch list + one full capture of the headline kernel.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -5 | tee gpurun_out/r02_pytest.txt
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras --sustain-s 0 > gpurun_out/r02_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:alz_biquad_tma -c 1 -o gpurun_out/r02_slaney python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras --sustain-s 0 > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_slaney.ncu-rep --page raw --csv > gpurun_out/r02_slaney_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
