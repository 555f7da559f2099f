timeout 900 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -2
echo "=== op profile"; timeout 300 python scripts/op_profile.py --batch 128 --top 12 2>&1 | cut -c1-250 | grep "total backbone\|fc1"
echo "=== bench bf16 B=256"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1800
