#!/bin/bash
# round 2, GPU call 26: FFMA2 SiLU epilogue in tc_conv_kernel, pair GEMMs for long K only: tests, profile, bench
mkdir -p gpurun_out
O=gpurun_out/r2_26
timeout 900 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_tc.py tests/test_gpu_fmb.py -x -q -k "(bf16 and not tf32x3) or tc_ops or fused" > ${O}_tests.log 2>&1; rc=$?
tail -2 ${O}_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|rel err|timed out" ${O}_tests.log | head -20 | cut -c1-300; exit 0; fi
timeout 120 python scripts/op_profile.py --batch 256 --top 40 2>&1 | cut -c1-200 > ${O}_op_profile.txt
head -1 ${O}_op_profile.txt | cut -c1-420; grep -E "tc_conv_kernel|fmb" ${O}_op_profile.txt | head -14 | cut -c1-120
timeout 600 python bench.py --no-cpu-baseline --no-frames --no-parity-line 2> ${O}_bench.err | tail -1 > ${O}_bench.json; cut -c1-300 ${O}_bench.json
