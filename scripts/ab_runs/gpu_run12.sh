for sh in 32x32 64x256 224x1344 1344x224; do
  echo "=== trace $sh"; MTB_TC_TRACE=$sh timeout 300 python scripts/op_profile.py --batch 128 --top 3 2>&1 | grep -A4 "MTB_TC_TRACE" | cut -c1-700
done
echo "=== dw check"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -k "fused_depthwise or (tc_ops and v2-s)" 2>&1 | grep -v "^$" | tail -5
timeout 300 python scripts/op_profile.py --batch 128 --top 8 2>&1 | tail -11
