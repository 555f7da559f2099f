#!/bin/bash
# round 2, GPU call 21: CTA-pair (cta_group::2) fmb_kernel: parity tests, then A/B against the single-CTA kernel
mkdir -p gpurun_out
O=gpurun_out/r2_21
timeout 300 python -m pytest tests/test_gpu_fmb.py -x -q -s > ${O}_fmb_tests.log 2>&1; rc=$?
grep -E "fused vs|passed|failed|Error|error|assert|timed out" ${O}_fmb_tests.log | cut -c1-220 | head -20
if [ $rc -ne 0 ]; then tail -30 ${O}_fmb_tests.log | cut -c1-300; exit 0; fi
for pr in 1 0; do
  MTB_FMB_PAIR=$pr timeout 120 python scripts/op_profile.py --batch 256 --top 8 2>&1 | grep -E "fmb_kernel" | cut -c1-130 | sed "s/^/pair=$pr /" | tee -a ${O}_ab.txt
done
MTB_FMB_TRACE=64 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace64.txt
timeout 600 python bench.py --no-cpu-baseline --no-frames --no-parity-line 2> ${O}_bench.err | tail -1 > ${O}_bench.json; cut -c1-300 ${O}_bench.json
