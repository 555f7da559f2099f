#!/bin/bash
# round 2, GPU call 20: ncu --set full of fmb_kernel (both block shapes) and of the direct-load scaled projection GEMM;
# the reports are exported to CSV on the box (raw page + SASS source page) - three .ncu-rep files exceed the 64 MiB return limit
mkdir -p gpurun_out
O=gpurun_out/r2_20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmb_kernel -c 2 -f -o /tmp/fmb python scripts/ncu_ops.py --precision bf16 --batch 256 --fused --ops backbone.1.2.1.block.0,backbone.1.3.1.block.0 > ${O}_ncu1.log 2>&1
MTB_FUSE_SE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 1 -f -o /tmp/proj_fused python scripts/ncu_ops.py --precision bf16 --batch 256 --ops backbone.1.6.1.block.3 > ${O}_ncu2.log 2>&1
MTB_FUSE_SE=0 timeout 600 ncu --set full --clock-control none -k regex:tc_conv_kernel -c 1 -f -o /tmp/proj_unfused python scripts/ncu_ops.py --precision bf16 --batch 256 --ops backbone.1.6.1.block.3 > ${O}_ncu3.log 2>&1
for r in fmb proj_fused proj_unfused; do
  ncu -i /tmp/$r.ncu-rep --page raw --csv > ${O}_$r.raw.csv 2>/dev/null
  python scripts/ncu_summary.py /tmp/$r.ncu-rep ${O}_$r.summary.csv
done
ncu -i /tmp/proj_fused.ncu-rep --page source --csv --print-source sass > ${O}_proj_fused.source.csv 2>/dev/null
ncu -i /tmp/fmb.ncu-rep --page source --csv --print-source sass --kernel-id :::1 > ${O}_fmb.source.csv 2>/dev/null
ls -la gpurun_out/ | grep r2_20 | awk '{print $5, $9}'
