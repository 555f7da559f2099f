#!/bin/bash
# round 2, GPU call 18: squeeze-excitation tail in one launch (se_fc2_scale_kernel): tests, profile, bench
mkdir -p gpurun_out
O=gpurun_out/r2_18
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q > ${O}_tests.log 2>&1; rc=$?
tail -2 ${O}_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|rel err" ${O}_tests.log | head -20 | cut -c1-300; exit 0; fi
for f in 1 0; do
  MTB_SE_APPLY=$f timeout 120 python scripts/op_profile.py --batch 256 --top 30 2>&1 | cut -c1-250 > ${O}_op_profile_seapply$f.txt
  head -1 ${O}_op_profile_seapply$f.txt | cut -c1-520; grep -E "fc1|fc2" ${O}_op_profile_seapply$f.txt | cut -c1-130
done
timeout 600 python bench.py --no-cpu-baseline --no-frames --no-parity-line 2> ${O}_bench.err | tail -1 > ${O}_bench.json; cut -c1-300 ${O}_bench.json
