#!/bin/bash
# round 2, GPU call 25: CTA-pair flat GEMMs (tc_conv_kernel PAIRM): parity vs conv2d at bench batch, A/B, bench
mkdir -p gpurun_out
O=gpurun_out/r2_25
timeout 900 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_tc.py -x -q -k "(bf16 and not tf32x3) or tc_ops" > ${O}_tests.log 2>&1; rc=$?
tail -2 ${O}_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|rel err|timed out" ${O}_tests.log | head -20 | cut -c1-300; exit 0; fi
for pr in 1 0; do
  MTB_TC_PAIR=$pr timeout 120 python scripts/op_profile.py --batch 256 --top 30 2>&1 | cut -c1-160 > ${O}_op_profile_pair$pr.txt
  head -1 ${O}_op_profile_pair$pr.txt | cut -c1-330; grep -E "tc_conv_kernel" ${O}_op_profile_pair$pr.txt | head -12 | cut -c1-120
done
timeout 600 python bench.py --no-cpu-baseline --no-frames --no-parity-line 2> ${O}_bench.err | tail -1 > ${O}_bench.json; cut -c1-300 ${O}_bench.json
