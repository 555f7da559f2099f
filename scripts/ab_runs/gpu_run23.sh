mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -3
echo "=== trace 32x32"; MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 16 2>&1 | cut -c1-330 | tail -24
echo "=== bench bf16 B=256"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1800
