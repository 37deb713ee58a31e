#!/bin/bash
# 2-GPU checks: NCCL parity tool, bench at N=2 (default + channels), then sustained power/clock per arithmetic variant on GPU 0.
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/nccl_check.py > gpurun_out/r02_nccl_check_n$N.txt 2> gpurun_out/r02_nccl_check_n$N.err
tail -3 gpurun_out/r02_nccl_check_n$N.txt; tail -5 gpurun_out/r02_nccl_check_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 600 gpurun_out/r02_bench_n$N.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --sharding channels > gpurun_out/r02_bench_channels_n$N.json 2> gpurun_out/r02_bench_channels_n$N.err
tail -c 600 gpurun_out/r02_bench_channels_n$N.err
if [ "$N" = "2" ]; then
for v in "ALZ_NO_FP32_TIER=1" "ALZ_X=1" "ALZ_TIER_TOL=1e9"; do
  echo "== $v"; env $v python bench.py --no-e2e --no-cpu --no-extras --sustain-s 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('burst ms %.3f frac %.3f clocks %s' % (d['ms_per_step'], r['frac'], d['clocks']))
print('sustained ms %.3f frac %.3f clocks %s' % (r['sustained']['ms_per_step'], r['sustained']['frac'], r['sustained']['clocks']))"
done 2>&1 | tee gpurun_out/r02_power.txt
fi
