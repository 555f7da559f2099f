# GPU check of the tensor-core path: fp32 parity first, then every tcgen05 test in its own process (a trap poisons the
# CUDA context), then a bf16 bench line.
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
for k in "fused_head" "fused_depthwise" "tc_ops and tiny" "tc_ops and v2-s" "tc_ops and v2-l" "bf16_forward"; do
  echo "=== $k"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -k "$k" 2>&1 | grep -v "^$" | tail -25
done
echo "=== bench bf16"; timeout 900 python bench.py --precision bf16 --steps 10 --warmup 3 --batch 128 2>&1 | tail -12
