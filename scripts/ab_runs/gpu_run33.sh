run() { echo "=== $*"; env "$@" MTB_TC_TRACE_CTA=999 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 1 2>&1 | grep "per-CTA" | cut -c1-330; }
run MTB_TC_DEBUG=32 MTB_TC_GRID=74
run MTB_TC_DEBUG=32 MTB_TC_GRID=16
run MTB_TC_DEBUG=111
run MTB_TC_DEBUG=175
run MTB_TC_DEBUG=239
run MTB_TC_DEBUG=239 MTB_TC_GRID=16
