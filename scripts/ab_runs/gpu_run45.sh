# SE cluster kernel with per-cluster rotation, serpentine depthwise order, SE fusion back to opt-in
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
echo "=== op profile B=256 (default)"; timeout 300 python scripts/op_profile.py --batch 256 --top 24 2>&1 | cut -c1-330 | tee gpurun_out/op_profile_b256_r45.txt | head -28
echo "=== MTB_DW_REV=1"; MTB_DW_REV=1 timeout 300 python scripts/op_profile.py --batch 256 --top 24 2>&1 | cut -c1-330 | grep "total\|block.3\|block.1 "
echo "=== MTB_SE_CLUSTER=0"; MTB_SE_CLUSTER=0 timeout 300 python scripts/op_profile.py --batch 256 --top 24 2>&1 | cut -c1-330 | grep "total"
echo "=== bench bf16 B=256"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r45.json | cut -c1-600
echo "=== bench bf16 B=256 MTB_DW_REV=1"; MTB_DW_REV=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
