# residual prefetch ahead of the accumulator wait + scoped PDL for the SE chain
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
prof() { echo "=== $*"; env "$@" timeout 300 python scripts/op_profile.py --batch 256 --top 30 2>&1 | cut -c1-330 | grep "total\|1.1.0.block.0\|block.3 \|2.1.block.1\|3.1.block.1" | cut -c1-420; }
prof A=1
prof MTB_PDL_SE=1
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
echo "=== bench MTB_PDL_SE=1"; MTB_PDL_SE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
