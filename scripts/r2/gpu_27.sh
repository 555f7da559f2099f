#!/bin/bash
# round 2, GPU call 27: specialised stem kernel: bit-equality + profile + bench
mkdir -p gpurun_out
O=gpurun_out/r2_27
timeout 600 python -m pytest tests/test_gpu_stem.py tests/test_gpu_parity.py -x -q -s -k "stem or tiny_model" > ${O}_tests.log 2>&1; rc=$?
grep -E "stem vs|passed|failed" ${O}_tests.log | cut -c1-200
if [ $rc -ne 0 ]; then grep -E "Error|error|assert" ${O}_tests.log | head -20 | cut -c1-300; exit 0; fi
timeout 120 python scripts/op_profile.py --batch 256 --top 40 2>&1 | cut -c1-200 > ${O}_op_profile.txt
head -1 ${O}_op_profile.txt | cut -c1-300; grep -E "backbone.1.0 " ${O}_op_profile.txt | cut -c1-120
timeout 600 python bench.py --no-cpu-baseline --no-frames --no-parity-line 2> ${O}_bench.err | tail -1 > ${O}_bench.json; cut -c1-200 ${O}_bench.json
