for m in plain host prof host_close prof_close; do echo "--- repro $m"; MTB_TRACE_DESTROY=1 timeout 120 python scripts/repro_destroy.py $m 2>&1 | tail -8; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for k in "fused_depthwise" "tc_ops and tiny" "tc_ops and v2-s" "tc_ops and v2-l" "bf16_forward"; do
  echo "=== $k"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -k "$k" 2>&1 | grep -v "^$" | tail -6
done
timeout 300 python scripts/op_profile.py --batch 128 --top 26 2>&1 | tail -30
echo "=== bench bf16 B=128"; timeout 900 python bench.py --precision bf16 --steps 10 --warmup 3 --batch 128 --no-cpu-baseline 2>&1 | grep '^{' | cut -c1-1800
echo "=== bench bf16 B=256"; timeout 900 python bench.py --precision bf16 --steps 10 --warmup 3 --batch 256 --no-cpu-baseline 2>&1 | grep '^{' | cut -c1-400
