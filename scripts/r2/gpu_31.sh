#!/bin/bash
# round 2, GPU call 31: fp32 variant of the TMA-staged depthwise kernel in the 3xTF32 parity mode: parity tests, A/B, bench
mkdir -p gpurun_out
O=gpurun_out/r2_31
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q -k "parity_modes or depthwise or tiny_model" > ${O}_tests.log 2>&1; rc=$?
tail -2 ${O}_tests.log | cut -c1-200
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|rel err|timed out" ${O}_tests.log | head -12 | cut -c1-300; exit 0; fi
for v in 1 0; do
  MTB_DW_F32_TMA=$v timeout 200 python bench.py --precision tf32x3 --steps 5 --warmup 3 --no-cpu-baseline --no-frames 2>/dev/null | tail -1 > ${O}_bench_tf32x3_dwtma$v.json
  python - <<PY
import json
d=json.load(open('${O}_bench_tf32x3_dwtma$v.json'))
print('dw_f32_tma=$v', round(d['value']), 'crops/s', round(d['ms_per_step'],2), 'ms parity', d['parity']['joints_rel_err_vs_oracle'], {k:v for k,v in d['roofline']['class_ms_warm_step'].items() if 'dw' in k})
PY
done
