# per-op table + destroy trace + ncu of the tensor-core conv kernel + the TF-only backbone parity tests
mkdir -p gpurun_out
timeout 300 python scripts/op_profile.py --batch 128 2>&1 | tail -60
echo "=== destroy trace"; MTB_TRACE_DESTROY=1 timeout 300 python bench.py --size s --batch 16 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -12
echo "=== parity (fp32, incl. TF-only backbones)"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s 2>&1 | tail -12
echo "=== ncu"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 30 -c 6 -o gpurun_out/tc_conv_r1 python scripts/op_profile.py --batch 64 > gpurun_out/ncu_tc_conv.log 2>&1; tail -3 gpurun_out/ncu_tc_conv.log
