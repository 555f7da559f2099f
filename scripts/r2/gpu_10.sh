#!/bin/bash
# round 2, GPU call 10 (re-entry baseline): all gpu tests, default bench line, per-op profiles of both tensor-core modes
mkdir -p gpurun_out
O=gpurun_out/r2_10
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > ${O}_gpu_tests.log 2>&1; tail -5 ${O}_gpu_tests.log
( time timeout 900 python bench.py ) 2> ${O}_bench.err | tail -1 > ${O}_bench_default.json; cut -c1-1500 ${O}_bench_default.json; tail -4 ${O}_bench.err
timeout 300 python scripts/op_profile.py --batch 256 --top 45 2>&1 | cut -c1-250 > ${O}_op_profile_bf16_b256.txt
timeout 300 python scripts/op_profile.py --batch 256 --precision tf32x3 --top 45 2>&1 | cut -c1-250 > ${O}_op_profile_tf32x3_b256.txt
head -3 ${O}_op_profile_bf16_b256.txt | cut -c1-400; head -3 ${O}_op_profile_tf32x3_b256.txt | cut -c1-400
