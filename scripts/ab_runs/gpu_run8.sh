timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for k in "fused_depthwise" "tc_ops and tiny" "tc_ops and v2-s" "tc_ops and v2-l"; do
  echo "=== $k"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -s -k "$k" 2>&1 | grep -v "^$" | tail -6
done
timeout 300 python scripts/op_profile.py --batch 128 --top 24 2>&1 | tail -28
echo "=== bench bf16 B=256"; MTB_TRACE_DESTROY=1 timeout 900 python bench.py --precision bf16 --steps 10 --warmup 3 --batch 256 --no-cpu-baseline 2>&1 | cut -c1-700 | tail -12
echo "=== ncu stage-1 convs"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 0 -c 3 -o gpurun_out/tc_conv_stage1_r1 python scripts/op_profile.py --batch 64 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
