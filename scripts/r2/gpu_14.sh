#!/bin/bash
# round 2, GPU call 14: fmb_kernel v2 (two weight producers, one issue block per kernel row, FFMA2 epilogue): tests, profile, trace
mkdir -p gpurun_out
O=gpurun_out/r2_14
timeout 600 python -m pytest tests/test_gpu_fmb.py -x -q -s > ${O}_fmb_tests.log 2>&1; rc=$?
grep -E "fused vs|passed|failed|Error|error|assert" ${O}_fmb_tests.log | cut -c1-220 | head -20
if [ $rc -ne 0 ]; then tail -30 ${O}_fmb_tests.log | cut -c1-300; exit 0; fi
for d in 0 2 32; do
  MTB_FMB_DEBUG=$d timeout 120 python scripts/op_profile.py --batch 256 --top 8 2>&1 | grep -E "fmb_kernel" | cut -c1-130 | sed "s/^/debug=$d /" | tee -a ${O}_ablation.txt
done
MTB_FMB_TRACE=64 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace64.txt
MTB_FMB_TRACE=96 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace96.txt
