#!/bin/bash
# round 2, GPU call 29: fmb_kernel with two A2 buffers (MTB_FMB_NA2=2) vs one
mkdir -p gpurun_out
O=gpurun_out/r2_29
for a in 1 2; do
  MTB_FMB_NA2=$a timeout 120 python scripts/op_profile.py --batch 256 --top 12 2>&1 | grep -E "fmb_kernel" | cut -c1-130 | sed "s/^/na2=$a /" | tee -a ${O}_ab.txt
done
MTB_FMB_NA2=2 timeout 200 python -m pytest tests/test_gpu_fmb.py -x -q -k "bench_batch or l-256 or s-256" 2>&1 | tail -1
