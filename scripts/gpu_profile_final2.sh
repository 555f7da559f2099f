# round-1 final evidence run: all gpu tests, smoke, bench lines (default + reference arm + other configs), head sweep,
# ncu launch list (+DRAM bytes) of the bench command, ncu --set full of the top kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench default (bf16, EffNetV2-L@256, 256 crops)"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_bf16.json | cut -c1-2600
echo "=== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-700
echo "=== bench fp32 parity mode"; timeout 600 python bench.py --precision fp32 --batch 64 --steps 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32.json | cut -c1-330
echo "=== bench c2 ResNet-50 stride 8, D=32, J=24, 128 crops"; timeout 600 python bench.py --size resnet50 --stride 8 --depth 32 --batch 128 --steps 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2_resnet50.json | cut -c1-330
echo "=== bench c4 EffNetV2-S J=122, 64 crops"; timeout 600 python bench.py --size s --joints 122 --batch 64 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c4.json | cut -c1-330
echo "=== bench c3 EffNetV2-L@384, 32 crops/GPU"; timeout 600 python bench.py --side 384 --batch 32 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c3_384.json | cut -c1-330
timeout 600 python scripts/head_sweep.py > gpurun_out/head_sweep.jsonl 2> gpurun_out/head_sweep.err; tail -1 gpurun_out/head_sweep.jsonl | cut -c1-200
echo "=== op profile"; timeout 300 python scripts/op_profile.py --batch 256 --top 45 2>&1 | cut -c1-250 > gpurun_out/op_profile_final.txt; head -2 gpurun_out/op_profile_final.txt | cut -c1-400
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1450 -c 465 --csv --log-file gpurun_out/bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; wc -l gpurun_out/bench_launches.csv
echo "=== ncu full: tensor-core conv/GEMM kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 60 -c 3 -o gpurun_out/tc_conv_r1_final python scripts/op_profile.py --batch 128 > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
echo "=== ncu full: TMA depthwise kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw3x3s1_tma -s 15 -c 1 -o gpurun_out/r1_dw_tma_v2 python scripts/op_profile.py --batch 256 > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
echo "=== ncu full: standalone soft-argmax + fused head"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"softargmax_bdjhw|tc_head_kernel" -s 30 -c 3 -o gpurun_out/softargmax_r1 python scripts/head_sweep.py > gpurun_out/ncu5.log 2>&1; tail -1 gpurun_out/ncu5.log
# keep the merge-back small: export the raw metric pages here, drop the reports (the tensor-core one is kept for the source view)
for r in r1_dw_tma_v2 softargmax_r1; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null && rm -f gpurun_out/$r.ncu-rep
done
ncu -i gpurun_out/tc_conv_r1_final.ncu-rep --page raw --csv > gpurun_out/tc_conv_r1_final.raw.csv 2>/dev/null
ls -la gpurun_out/ | awk '{print $5, $9}' | tail -30
echo "=== A/B: MTB_PDL_SE=1"; MTB_PDL_SE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
echo "=== A/B: MTB_ENABLE_PDL=1"; MTB_ENABLE_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
echo "=== A/B: --graph 1"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --graph 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['cuda_graph_replay_crops_per_s'])"
