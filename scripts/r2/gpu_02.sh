#!/bin/bash
# round 2, GPU call 2: 3xTF32 with out-of-tensor-core accumulation (chain drains), multiperson kernels, MTB_GRAPH fix
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=20 > gpurun_out/r2_02_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/r2_02_gpu_tests.log
for chain in 1 2 4; do
  MTB_T32_CHAIN=$chain timeout 600 python bench.py --precision tf32x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_02_bench_tf32x3_chain$chain.json 2>> gpurun_out/r2_02_bench.err
done
MTB_T32_CHAIN=4 timeout 300 python -m pytest tests/test_gpu_tf32.py -q -s -k "vs_conv2d" > gpurun_out/r2_02_tf32_chain4.log 2>&1
MTB_GRAPH=1 timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_02_bench_bf16_graph.json 2>> gpurun_out/r2_02_bench.err
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_02_bench_bf16.json 2>> gpurun_out/r2_02_bench.err
timeout 600 python scripts/op_profile.py --precision tf32x3 > gpurun_out/r2_02_op_profile_tf32x3.txt 2>&1
grep -E "passed|failed" gpurun_out/r2_02_gpu_tests.log | tail -3
