#!/bin/bash
# round 2, GPU call 15: SE scale fused into the projection GEMM (8 scaler warps) vs the separate se_scale_kernel pass
mkdir -p gpurun_out
O=gpurun_out/r2_15
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tf32.py -x -q -k "tc_ops or (bf16 and not tf32x3)" > ${O}_tests.log 2>&1; rc=$?
tail -3 ${O}_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|rel err" ${O}_tests.log | head -20 | cut -c1-300; exit 0; fi
for f in 1 0; do
  MTB_FUSE_SE=$f timeout 120 python scripts/op_profile.py --batch 256 --top 12 2>&1 | cut -c1-250 > ${O}_op_profile_fuse$f.txt
  head -1 ${O}_op_profile_fuse$f.txt | cut -c1-420; grep -E "block.3 " ${O}_op_profile_fuse$f.txt | cut -c1-130
done
timeout 600 python bench.py --no-cpu-baseline --no-frames --no-parity-line 2> ${O}_bench.err | tail -1 > ${O}_bench.json; cut -c1-300 ${O}_bench.json
