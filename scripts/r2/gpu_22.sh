#!/bin/bash
# round 2, GPU call 22: fmb pair kernel, epilogue-2 timeline
mkdir -p gpurun_out
O=gpurun_out/r2_22
timeout 300 python -m pytest tests/test_gpu_fmb.py -x -q -k "bench_batch or l-256" > ${O}_fmb_tests.log 2>&1; tail -1 ${O}_fmb_tests.log
MTB_FMB_PAIR=1 timeout 120 python scripts/op_profile.py --batch 256 --top 12 2>&1 | grep -E "fmb_kernel" | cut -c1-130 | tee -a ${O}_ab.txt
MTB_FMB_TRACE=64 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace64.txt
MTB_FMB_TRACE=96 timeout 120 python scripts/op_profile.py --batch 256 --top 3 > /dev/null 2> ${O}_trace96.txt
