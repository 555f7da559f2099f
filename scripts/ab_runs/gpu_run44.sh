# fused SE scalers (4 warps), SE cluster kernel, FFMA2 depthwise: correctness + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== op profile B=256 (all new)"; timeout 300 python scripts/op_profile.py --batch 256 --top 22 2>&1 | cut -c1-200 | tee gpurun_out/op_profile_b256_r44.txt | head -26
echo "=== MTB_FUSE_SE=0"; MTB_FUSE_SE=0 timeout 300 python scripts/op_profile.py --batch 256 --top 8 2>&1 | cut -c1-330 | grep "block.3\|total"
echo "=== MTB_SE_CLUSTER=0"; MTB_SE_CLUSTER=0 timeout 300 python scripts/op_profile.py --batch 256 --top 12 2>&1 | cut -c1-330 | grep "fc1\|total"
echo "=== bench bf16 B=256"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r44.json | cut -c1-1500
