#!/bin/bash
# round 2, GPU call 12: fmb_kernel ablations (what bounds a tile: weight stream / epilogue / patch / MMA)
mkdir -p gpurun_out
O=gpurun_out/r2_12
for d in 0 2 8 16 32 10 42 58; do
  MTB_FMB_DEBUG=$d timeout 120 python scripts/op_profile.py --batch 256 --top 6 2>&1 | grep -E "fmb_kernel" | cut -c1-130 | sed "s/^/debug=$d /" | tee -a ${O}_ablation.txt
done
