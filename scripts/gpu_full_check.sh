# full GPU check: every gpu test (parity in one process, tensor-core tests file-by-file), smoke, bench lines
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_tc.py -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench bf16 B=256 (PDL on)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-1500 | tail -2
echo "=== bench bf16 B=256 (PDL off)"; MTB_DISABLE_PDL=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-200 | tail -2
