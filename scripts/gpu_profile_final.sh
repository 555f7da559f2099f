# evidence run: tests, bench lines, head / soft-argmax sweep, ncu launch list (+DRAM bytes) of the bench command, ncu --set full
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q 2>&1 | tail -3
echo "=== bench default (bf16, EffNetV2-L@256, 256 crops)"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_bf16.json | cut -c1-2500
echo "=== bench fp32 parity mode"; timeout 900 python bench.py --precision fp32 --batch 64 --steps 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32.json | cut -c1-400
echo "=== bench c2 ResNet-50 stride 8, D=32, J=24, 128 crops"; timeout 900 python bench.py --size resnet50 --stride 8 --depth 32 --batch 128 --steps 5 --cpu-sample 4 2>&1 | tail -1 | tee gpurun_out/bench_c2_resnet50.json | cut -c1-1200
echo "=== bench c4 EffNetV2-S J=122, 64 crops"; timeout 900 python bench.py --size s --joints 122 --batch 64 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c4.json | cut -c1-400
echo "=== bench c3 EffNetV2-L@384, 32 crops/GPU"; timeout 900 python bench.py --side 384 --batch 32 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c3_384.json | cut -c1-400
timeout 600 python scripts/head_sweep.py > gpurun_out/head_sweep.jsonl 2> gpurun_out/head_sweep.err; tail -2 gpurun_out/head_sweep.jsonl | cut -c1-200
echo "=== ncu launch list (bench.py --steps 2 --warmup 3 --batch 256)"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1950 -c 660 --csv --log-file gpurun_out/bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/bench_launches.csv
echo "=== ncu full: tensor-core conv/GEMM kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 60 -c 5 -o gpurun_out/tc_conv_r1_final python scripts/op_profile.py --batch 128 > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
echo "=== ncu full: standalone soft-argmax + fused head"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"softargmax_bdjhw|tc_head_kernel" -s 30 -c 3 -o gpurun_out/softargmax_r1 python scripts/head_sweep.py > gpurun_out/ncu5.log 2>&1; tail -1 gpurun_out/ncu5.log
