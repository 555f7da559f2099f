# chunking correctness + A/B, in-kernel trace of the stage-5 projection GEMM, ncu full of the TMA depthwise + projection kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "chunked" 2>&1 | tail -3
run_bench() { echo "=== bench MTB_CHUNKS=$1"; MTB_CHUNKS=$1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['class_ms_first_step'])"; }
run_bench 0,0,0,0,0,0,0
run_bench 0,0,0,64,64,128,128
run_bench 0,0,0,32,32,64,64
run_bench 64,32,32,64,64,128,128
echo "=== trace 1344x224 B=256"; MTB_TC_TRACE=1344x224 timeout 300 python scripts/op_profile.py --batch 256 --top 1 2>&1 | grep -A6 "per-CTA totals\|MTB_TC_TRACE" | cut -c1-1500
echo "=== trace 224x1344 B=256"; MTB_TC_TRACE=224x1344 timeout 300 python scripts/op_profile.py --batch 256 --top 1 2>&1 | grep -A6 "per-CTA totals\|MTB_TC_TRACE" | cut -c1-1500
echo "=== ncu dw"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw3x3s1_tma -s 15 -c 1 -o gpurun_out/r1_dw_tma python scripts/op_profile.py --batch 256 --top 1 > gpurun_out/ncu_dw.log 2>&1; tail -2 gpurun_out/ncu_dw.log
echo "=== ncu project"; timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_conv_kernel<\D*0, \D*1, \D*64' -s 26 -c 1 -o gpurun_out/r1_tc_project python scripts/op_profile.py --batch 256 --top 1 > gpurun_out/ncu_proj.log 2>&1; tail -2 gpurun_out/ncu_proj.log
ls -la gpurun_out/*.ncu-rep
