#!/bin/bash
# round 2, GPU call 3: tf32x3 splitter/epilogue optimisations + fp32 strip depthwise, multiperson test fixes, new bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=20 > gpurun_out/r2_03_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/r2_03_gpu_tests.log
for chain in 2 4; do
  MTB_T32_CHAIN=$chain timeout 600 python bench.py --precision tf32x3 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_03_bench_tf32x3_chain$chain.json 2>> gpurun_out/r2_03_bench.err
done
timeout 600 python scripts/op_profile.py --precision tf32x3 > gpurun_out/r2_03_op_profile_tf32x3.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2_03_bench_default.json 2>> gpurun_out/r2_03_bench.err
MTB_GRAPH=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-parity-line > gpurun_out/r2_03_bench_bf16_graph.json 2>> gpurun_out/r2_03_bench.err
grep -E "passed|failed" gpurun_out/r2_03_gpu_tests.log | tail -3
