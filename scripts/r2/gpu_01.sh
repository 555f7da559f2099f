#!/bin/bash
# round 2, GPU call 1: first run of the 3xTF32 kernel - per-op tests vs conv2d, full-model parity, first bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_01_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tf32.py -q --maxfail=6 -s > gpurun_out/r2_01_tf32_tests.log 2>&1
echo "tf32 tests exit $?" >> gpurun_out/r2_01_tf32_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "parity_modes or tiny_model" -s > gpurun_out/r2_01_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/r2_01_parity.log
timeout 600 python bench.py --precision tf32x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_01_bench_tf32x3.json 2> gpurun_out/r2_01_bench_tf32x3.err
timeout 600 python bench.py --precision tf32x3 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_01_bench_tf32x3_b64.json 2>> gpurun_out/r2_01_bench_tf32x3.err
MTB_T32_RB=128 timeout 600 python bench.py --precision tf32x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_01_bench_tf32x3_rb128.json 2>> gpurun_out/r2_01_bench_tf32x3.err
MTB_GRAPH=1 timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_01_bench_bf16_graph.json 2> gpurun_out/r2_01_bench_bf16_graph.err
timeout 600 python scripts/op_profile.py --precision tf32x3 > gpurun_out/r2_01_op_profile_tf32x3.txt 2>&1
tail -5 gpurun_out/r2_01_tf32_tests.log gpurun_out/r2_01_parity.log
