# microbenchmark + validation of the pending edits (tmem_base shuffle, dw OW=2, graph replay)
mkdir -p gpurun_out
echo "=== mma_rate"; timeout 120 scripts/bin/mma_rate 2>&1 | tee gpurun_out/mma_rate.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_tc.py -q 2>&1 | tail -3
echo "=== op profile"; timeout 600 python scripts/op_profile.py --batch 128 --top 14 2>&1 | cut -c1-200 | tail -18
echo "=== bench bf16 B=256"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1800
echo "=== bench bf16 B=256 graph"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --graph 1 2>&1 | tail -1 | cut -c1-900
