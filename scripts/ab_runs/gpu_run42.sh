# K rotation + TMA depthwise: correctness, then A/B per-op profile, then bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
echo "=== op profile B=256 (rot + dw tma)"; timeout 300 python scripts/op_profile.py --batch 256 --top 24 2>&1 | cut -c1-200 | tee gpurun_out/op_profile_b256_new.txt | head -28
echo "=== op profile B=256 MTB_TC_ROT=0"; MTB_TC_ROT=0 timeout 300 python scripts/op_profile.py --batch 256 --top 12 2>&1 | cut -c1-200 | grep "block.3\|total"
echo "=== bench bf16 B=256"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1800
