#!/bin/bash
# round 2, GPU call 9: coalesced column-domain epilogue of the 3xTF32 kernel; fp16 / packed 16-bit soft-argmax
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_parity.py -q -s -k "tf32x3 or parity_modes or soft_argmax" > gpurun_out/r2_09_tests.log 2>&1
timeout 300 python scripts/op_profile.py --precision tf32x3 --top 22 > gpurun_out/r2_09_op_profile.txt 2>&1
MTB_T32_DEBUG=4 timeout 300 python scripts/op_profile.py --precision tf32x3 --top 8 > gpurun_out/r2_09_op_profile_noepi.txt 2>&1
timeout 300 python bench.py --precision tf32x3 --steps 5 --warmup 3 --no-cpu-baseline --no-frames > gpurun_out/r2_09_bench_tf32x3.json 2> gpurun_out/r2_09_bench.err
timeout 300 python scripts/head_sweep.py > gpurun_out/r2_09_head_sweep.jsonl 2> gpurun_out/r2_09_head_sweep.err
grep -E "passed|failed" gpurun_out/r2_09_tests.log
head -1 gpurun_out/r2_09_op_profile*.txt | cut -c1-300
grep softargmax gpurun_out/r2_09_head_sweep.jsonl | cut -c1-200
