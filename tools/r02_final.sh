#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -4 | tee gpurun_out/r02_pytest.txt
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_ref_n1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.txt
