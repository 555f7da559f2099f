#!/bin/bash
# round 2, GPU call 23: in-kernel timeline of the stage-1 conv (32->32 3x3 @128x128, mode 2) to find its per-tile bound
mkdir -p gpurun_out
O=gpurun_out/r2_23
MTB_TC_TRACE=32x32 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace_32x32.txt; head -c 6000 ${O}_trace_32x32.txt
for d in 0 1 4 8 64; do
  MTB_TC_DEBUG=$d timeout 120 python scripts/op_profile.py --batch 256 --top 12 2>&1 | grep -E "1.1.0.block.0" | cut -c1-110 | sed "s/^/debug=$d /" | tee -a ${O}_ablation.txt
done
