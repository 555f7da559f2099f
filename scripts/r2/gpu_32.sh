#!/bin/bash
# round 2, GPU call 32: last sanity of the committed build: smoke (all modes), bench default (short)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 300 python bench.py --no-frames --no-cpu-baseline --steps 5 2>/dev/null | tail -1 > gpurun_out/r2_32_bench.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2_32_bench.json'))
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'parity_mode', round(d['parity_mode']['value']), d['parity_mode']['parity']['joints_rel_err_vs_oracle'])
PY
