for dbg in 0 15; do
echo "=== trace 32x32 debug=$dbg"; MTB_TC_DEBUG=$dbg MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 3 2>&1 | grep -A4 "MTB_TC_TRACE" | cut -c1-900
done
for k in "tf_backbones"; do
  echo "=== $k"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -k "$k" 2>&1 | grep -v "^$" | tail -12
done
