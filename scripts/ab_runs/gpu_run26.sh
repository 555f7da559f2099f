echo "=== trace 32x32 debug=16"; MTB_TC_DEBUG=16 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 1 2>&1 | cut -c1-600 | head -4
echo "=== trace 32x32 debug=31"; MTB_TC_DEBUG=31 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 1 2>&1 | cut -c1-600 | head -4
