run() { echo "=== $*"; env "$@" MTB_TC_TRACE_CTA=999 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 1 2>&1 | grep "per-CTA" | cut -c1-330; }
run MTB_TC_DEBUG=288
run MTB_TC_DEBUG=303
run MTB_TC_DEBUG=32
