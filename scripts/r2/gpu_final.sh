#!/bin/bash
# round-2 evidence run (one B200): all gpu tests, smoke, the bench lines (default = bf16 headline + tf32x3 parity-mode sibling +
# parity + frames leg + CPU baseline; reference arm; tf32x3 / fp32 lines; configs c2 / c3 / c4), head sweep, ncu launch lists
# (+DRAM bytes) of the bench command in both modes, ncu --set full of the top kernels.  Outputs under gpurun_out/r2_final_*.
mkdir -p gpurun_out
O=gpurun_out/r2_final
timeout 1500 python -m pytest tests -m gpu -q > ${O}_gpu_tests.log 2>&1; tail -2 ${O}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -4 ${O}_smoke.log
echo "=== bench default"; timeout 900 python bench.py 2> ${O}_bench.err | tail -1 > ${O}_bench_default.json; cut -c1-300 ${O}_bench_default.json
echo "=== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>> ${O}_bench.err | tail -1 > ${O}_bench_reference.json; cut -c1-200 ${O}_bench_reference.json
echo "=== bench tf32x3 (parity mode as its own line)"; timeout 600 python bench.py --precision tf32x3 --steps 5 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_tf32x3.json
echo "=== bench fp32 (CUDA-core parity mode), 64 crops"; timeout 600 python bench.py --precision fp32 --batch 64 --steps 5 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_fp32.json
echo "=== c2 ResNet-50 stride 8, D=32, J=24, 128 crops"; timeout 600 python bench.py --size resnet50 --stride 8 --depth 32 --batch 128 --steps 5 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_c2_resnet50.json
echo "=== c4 EffNetV2-S J=122, 64 crops"; timeout 600 python bench.py --size s --joints 122 --batch 64 --steps 10 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_c4.json
echo "=== c3 EffNetV2-L@384, 32 crops (one GPU's share of 256)"; timeout 600 python bench.py --side 384 --batch 32 --steps 10 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_c3_384_b32.json
timeout 600 python scripts/head_sweep.py > ${O}_head_sweep.jsonl 2> ${O}_head_sweep.err; tail -1 ${O}_head_sweep.jsonl | cut -c1-200
timeout 300 python scripts/op_profile.py --batch 256 --top 45 2>&1 | cut -c1-250 > ${O}_op_profile_bf16_b256.txt
timeout 300 python scripts/op_profile.py --batch 256 --precision tf32x3 --top 45 2>&1 | cut -c1-250 > ${O}_op_profile_tf32x3_b256.txt
echo "=== ncu launch lists (one warm step each)"
MTB_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1400 -c 480 --csv --log-file ${O}_launches_bf16.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-frames --no-parity --no-parity-line > ${O}_under_ncu_bf16.log 2>&1; wc -l ${O}_launches_bf16.csv
MTB_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1100 -c 420 --csv --log-file ${O}_launches_tf32x3.csv \
  python bench.py --precision tf32x3 --steps 2 --warmup 3 --no-cpu-baseline --no-frames --no-parity > ${O}_under_ncu_tf32x3.log 2>&1; wc -l ${O}_launches_tf32x3.csv
echo "=== ncu full: tc_conv_kernel (bf16), tc32_conv_kernel (3xTF32), warp_crops_kernel"
OPS=backbone.1.5.1.block.3,backbone.1.5.1.block.0,backbone.1.2.1.block.0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 3 -f -o ${O}_tc_conv python scripts/ncu_ops.py --precision bf16 --batch 256 --ops $OPS > ${O}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc32_conv_kernel -c 3 -f -o ${O}_tc32_conv python scripts/ncu_ops.py --precision tf32x3 --batch 128 --ops $OPS > ${O}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"warp_crops_kernel|tta_merge_kernel|crop_setup_kernel|pyramid" -c 6 -f -o ${O}_multiperson python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-parity-line > ${O}_ncu3.log 2>&1
for r in tc_conv tc32_conv multiperson; do ncu -i ${O}_$r.ncu-rep --page raw --csv > ${O}_$r.raw.csv 2>/dev/null; done
rm -f ${O}_multiperson.ncu-rep
ls -la gpurun_out/ | grep r2_final | awk '{print $5, $9}'
