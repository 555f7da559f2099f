nvidia-smi --query-gpu=name,uuid,serial --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -2
echo "=== op profile"; MTB_TC_TRACE_CTA=999 MTB_TC_DEBUG=32 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 16 2>&1 | grep -v "producer\|mma(\|epilogue(\|per-CTA cycles" | cut -c1-200 | tail -20
echo "=== bench bf16 B=256"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1800
