echo "=== trace cta 0"; MTB_TC_DEBUG=32 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 1 2>&1 | grep "per-CTA" | cut -c1-1200
echo "=== trace cta 77"; MTB_TC_TRACE_CTA=77 MTB_TC_DEBUG=32 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 1 2>&1 | grep "per-CTA" | cut -c1-1200
echo "=== trace cta 999 (none)"; MTB_TC_TRACE_CTA=999 MTB_TC_DEBUG=32 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 1 2>&1 | grep "per-CTA" | cut -c1-1200
