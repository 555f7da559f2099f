mkdir -p gpurun_out
echo "=== trace 32x32 debug=0"; MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 8 2>&1 | cut -c1-420 | tail -14
echo "=== trace 32x32 debug=15"; MTB_TC_DEBUG=15 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 3 2>&1 | cut -c1-420 | head -5
echo "=== trace 32x32 debug=6 (no epilogue math/residual)"; MTB_TC_DEBUG=6 MTB_TC_TRACE=32x32 timeout 300 python scripts/op_profile.py --batch 128 --top 3 2>&1 | cut -c1-420 | head -5
echo "=== ncu stage 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 2 -o gpurun_out/tc_conv_stage1_v2 python scripts/op_profile.py --batch 64 > gpurun_out/ncu6.log 2>&1; tail -1 gpurun_out/ncu6.log
