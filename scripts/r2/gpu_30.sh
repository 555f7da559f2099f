#!/bin/bash
# round 2, GPU call 30: sanity after the host-side refactor of the fmb weight packing; bench default with traffic.json in place
mkdir -p gpurun_out
O=gpurun_out/r2_30
timeout 300 python -m pytest tests/test_gpu_fmb.py tests/test_gpu_stem.py -x -q > ${O}_tests.log 2>&1; tail -1 ${O}_tests.log
timeout 600 python bench.py --no-frames 2> ${O}_bench.err | tail -1 > ${O}_bench_default.json
python - <<PY
import json
d=json.load(open('${O}_bench_default.json'))
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'traffic', d['roofline']['traffic'], json.dumps(d['roofline'].get('tensor_core_kernels',{}).get('combined')))
print('parity_mode', round(d['parity_mode']['value']), d['parity_mode']['roofline'].get('traffic'))
PY
