#!/bin/bash
# round 2, GPU call 11: fused FusedMBConv kernel (fmb_kernel) - parity tests, then per-op profile and a bench line
mkdir -p gpurun_out
O=gpurun_out/r2_11
timeout 600 python -m pytest tests/test_gpu_fmb.py -x -q -s > ${O}_fmb_tests.log 2>&1; rc=$?
grep -E "fused vs|passed|failed|Error|error|rel err|assert" ${O}_fmb_tests.log | cut -c1-260 | head -40
if [ $rc -ne 0 ]; then tail -30 ${O}_fmb_tests.log | cut -c1-300; exit 0; fi
timeout 300 python scripts/op_profile.py --batch 256 --top 14 2>&1 | cut -c1-250 > ${O}_op_profile_bf16_b256.txt; head -16 ${O}_op_profile_bf16_b256.txt
MTB_FMB_NA2=2 timeout 300 python scripts/op_profile.py --batch 256 --top 6 2>&1 | cut -c1-250 > ${O}_op_profile_na2.txt; head -8 ${O}_op_profile_na2.txt
timeout 600 python bench.py --no-cpu-baseline --no-frames --no-parity-line 2> ${O}_bench.err | tail -1 > ${O}_bench.json; cut -c1-400 ${O}_bench.json
