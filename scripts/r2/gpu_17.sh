#!/bin/bash
# round 2, GPU call 17: where does the fused SE scaler lose its time (ablations)
mkdir -p gpurun_out
O=gpurun_out/r2_17
for d in 0 512 1024 1536; do
  MTB_TC_DEBUG=$d timeout 120 python scripts/op_profile.py --batch 256 --top 12 2>&1 | grep -E "block.3 " | cut -c1-110 | sed "s/^/debug=$d /" | tee -a ${O}_ablation.txt
done
MTB_TC_TRACE=2304x384 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace_2304x384.txt; head -c 3000 ${O}_trace_2304x384.txt
