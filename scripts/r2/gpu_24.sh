#!/bin/bash
# round 2, GPU call 24: one-chunk tiles: accumulator hand-back by the owning epilogue group only
mkdir -p gpurun_out
O=gpurun_out/r2_24
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tf32.py -x -q -k "tc_ops or (bf16 and not tf32x3)" > ${O}_tests.log 2>&1; rc=$?
tail -2 ${O}_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|rel err" ${O}_tests.log | head -20 | cut -c1-300; exit 0; fi
timeout 120 python scripts/op_profile.py --batch 256 --top 40 2>&1 | cut -c1-200 > ${O}_op_profile.txt; head -14 ${O}_op_profile.txt | cut -c1-150
timeout 600 python bench.py --no-cpu-baseline --no-frames --no-parity-line 2> ${O}_bench.err | tail -1 > ${O}_bench.json; cut -c1-300 ${O}_bench.json
