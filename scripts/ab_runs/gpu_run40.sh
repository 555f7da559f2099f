# re-entry baseline: all gpu tests, smoke, op profile, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== op profile"; timeout 300 python scripts/op_profile.py --batch 256 --top 45 2>&1 | cut -c1-250 | tee gpurun_out/op_profile_b256.txt | tail -50
echo "=== bench bf16 B=256"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_default.json | cut -c1-2500
