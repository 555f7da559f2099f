# evidence run: head / soft-argmax roofline sweep, ncu launch list of the bench command, ncu --set full of the top kernels
mkdir -p gpurun_out
timeout 600 python scripts/head_sweep.py > gpurun_out/head_sweep.jsonl 2> gpurun_out/head_sweep.err; tail -3 gpurun_out/head_sweep.jsonl | cut -c1-300
echo "=== ncu launch list (bench.py --steps 2 --warmup 3 --batch 256)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1950 -c 700 --csv --log-file gpurun_out/bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/bench_launches.csv
echo "=== ncu full: tensor-core conv/GEMM kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 60 -c 5 -o gpurun_out/tc_conv_r1_final python scripts/op_profile.py --batch 128 > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
echo "=== ncu full: depthwise + pool"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3 -s 20 -c 2 -o gpurun_out/dwconv_r1 python scripts/op_profile.py --batch 128 > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
echo "=== ncu full: standalone soft-argmax + fused head"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"softargmax_bdjhw|tc_head_kernel" -s 30 -c 3 -o gpurun_out/softargmax_r1 python scripts/head_sweep.py > gpurun_out/ncu5.log 2>&1; tail -1 gpurun_out/ncu5.log
