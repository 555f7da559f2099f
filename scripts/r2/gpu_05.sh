#!/bin/bash
# round 2, GPU call 5: ablation of the 3xTF32 kernel roles (MTB_T32_DEBUG: 1 no split, 2 no drain, 4 no epilogue), LDS/STS splitter
mkdir -p gpurun_out
for dbg in 0 1 2 4 7; do
  MTB_T32_DEBUG=$dbg timeout 300 python scripts/op_profile.py --precision tf32x3 --top 14 > gpurun_out/r2_05_op_profile_dbg$dbg.txt 2>&1
done
MTB_T32_CHAIN=1000 timeout 300 python scripts/op_profile.py --precision tf32x3 --top 14 > gpurun_out/r2_05_op_profile_chain1000.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multiperson.py tests/test_gpu_parity.py tests/test_gpu_tf32.py -q -s -k "crop_generation or checkpoint or tiny_model or vs_conv2d" > gpurun_out/r2_05_tests.log 2>&1
grep -E "passed|failed" gpurun_out/r2_05_tests.log
head -3 gpurun_out/r2_05_op_profile_dbg*.txt | cut -c1-250
