#!/bin/bash
# round 2, GPU call 13: in-kernel clock64 timeline of fmb_kernel (CTA 0), with and without the MMAs
mkdir -p gpurun_out
O=gpurun_out/r2_13
MTB_FMB_TRACE=64 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace64.txt
MTB_FMB_TRACE=64 MTB_FMB_DEBUG=58 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace64_dbg58.txt
MTB_FMB_TRACE=96 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace96.txt
wc -c ${O}_trace*.txt
