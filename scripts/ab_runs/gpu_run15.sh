for k in "tc_ops and tiny" "tc_ops and v2-s" "tc_ops and v2-l"; do
  echo "=== $k"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -k "$k" 2>&1 | grep -v "^$" | tail -4
done
timeout 300 python scripts/op_profile.py --batch 128 --top 14 2>&1 | tail -17
echo "=== bench bf16 B=256"; timeout 900 python bench.py --precision bf16 --steps 10 --warmup 3 --batch 256 --no-cpu-baseline 2>&1 | cut -c1-300 | tail -3
