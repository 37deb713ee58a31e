#!/bin/bash
# compute-sanitizer over the GPU tests that exercise the round-2 kernels (tiers, window, parallel sum, envelope, virtual streams, channel-major)
mkdir -p gpurun_out
SEL='precision_tiers or parallel_sum or fused_envelope or ragged_shapes or generic_kernel or wav_to_bank or channel_major or memory_zero or unaligned or block_split or lpc or callers_of_the_path'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --timeout 600 -k "$SEL" > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.txt
tail -6 gpurun_out/r02_sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --timeout 500 -k "time_parallel_batches and (33-40000 or 1000-20000)" > gpurun_out/r02_sanitizer_virtual.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_virtual.txt
tail -5 gpurun_out/r02_sanitizer_virtual.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --timeout 500 -k "parallel_sum or fused_envelope or channel_major" > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.txt
tail -5 gpurun_out/r02_sanitizer_racecheck.txt
