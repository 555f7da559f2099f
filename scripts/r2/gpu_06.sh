#!/bin/bash
# round 2, GPU call 6: 64-byte-row stages (6 in flight) vs 128; hi-rewrite skipped (MTB_T32_DEBUG=8) accuracy + speed
mkdir -p gpurun_out
timeout 300 python scripts/op_profile.py --precision tf32x3 --top 14 > gpurun_out/r2_06_op_profile_rb64.txt 2>&1
MTB_T32_RB=128 timeout 300 python scripts/op_profile.py --precision tf32x3 --top 14 > gpurun_out/r2_06_op_profile_rb128.txt 2>&1
MTB_T32_DEBUG=8 timeout 300 python scripts/op_profile.py --precision tf32x3 --top 14 > gpurun_out/r2_06_op_profile_rb64_nohi.txt 2>&1
MTB_T32_DEBUG=1 timeout 300 python scripts/op_profile.py --precision tf32x3 --top 14 > gpurun_out/r2_06_op_profile_rb64_nosplit.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_parity.py -q -s -k "tf32x3 or parity_modes" > gpurun_out/r2_06_tests_rb64.log 2>&1
MTB_T32_DEBUG=8 timeout 600 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_parity.py -q -s -k "tf32x3 or parity_modes" > gpurun_out/r2_06_tests_rb64_nohi.log 2>&1
timeout 300 python bench.py --precision tf32x3 --steps 5 --warmup 3 --no-cpu-baseline --no-frames > gpurun_out/r2_06_bench_tf32x3.json 2> gpurun_out/r2_06_bench.err
grep -E "passed|failed" gpurun_out/r2_06_tests_*.log
head -1 gpurun_out/r2_06_op_profile_*.txt | cut -c1-300
