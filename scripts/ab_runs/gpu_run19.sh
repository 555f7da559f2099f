timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_tc.py -q 2>&1 | tail -4
timeout 300 python scripts/op_profile.py --batch 128 --top 12 2>&1 | tail -15
timeout 600 python scripts/head_sweep.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'softargmax' in d['kernel']: print(d['kernel'], d['dtype'], 'B', d['B'], round(d['us'],1), 'us', round(d['GBps']), 'GB/s', round(100*d['frac_of_hbm_peak'],1), '% HBM')
    elif d['B'] in (256, 1024): print('head', d['reading'], 'B', d['B'], round(d['us'],1), 'us', round(d['TFLOPs'],1), 'TF/s')
"
echo "=== bench bf16 B=256"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-300 | tail -2
