#!/bin/bash
# round-2 final evidence run (one B200).  All gpu tests, smoke, bench lines (default = bf16 headline + tf32x3 parity-mode sibling
# + parity + frames leg + CPU baseline; reference arm; tf32x3 / fp32 lines; configs c2 / c3 / c4), head sweep, per-op profiles,
# ncu launch lists (+DRAM bytes) of the bench command in both modes, ncu --set full of the top kernels (exported to CSV on the box:
# the .ncu-rep files exceed the 64 MiB return limit).  Outputs under gpurun_out/r2_final_*.
mkdir -p gpurun_out
O=gpurun_out/r2_final
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_gpu_tests.log 2>&1; tail -4 ${O}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -5 ${O}_smoke.log
echo "=== bench default"; timeout 900 python bench.py 2> ${O}_bench.err | tail -1 > ${O}_bench_default.json; cut -c1-300 ${O}_bench_default.json
echo "=== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>> ${O}_bench.err | tail -1 > ${O}_bench_reference.json; cut -c1-200 ${O}_bench_reference.json
timeout 600 python bench.py --precision tf32x3 --steps 5 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_tf32x3.json
timeout 600 python bench.py --precision fp32 --batch 64 --steps 5 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_fp32.json
timeout 600 python bench.py --size resnet50 --stride 8 --depth 32 --batch 128 --steps 5 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_c2_resnet50.json
timeout 600 python bench.py --size s --joints 122 --batch 64 --steps 10 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_c4.json
timeout 600 python bench.py --side 384 --batch 32 --steps 10 --no-cpu-baseline --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_c3_384_b32.json
timeout 600 python bench.py --graph 1 --steps 10 --no-cpu-baseline --no-frames --no-parity --no-parity-line 2>> ${O}_bench.err | tail -1 > ${O}_bench_graph.json
for f in tf32x3 fp32 c2_resnet50 c4 c3_384_b32 graph; do python - <<PY
import json
d=json.load(open('${O}_bench_$f.json'))
print('$f', round(d['value']), 'crops/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), d.get('parity',{}).get('joints_rel_err_vs_oracle'))
PY
done
timeout 600 python scripts/head_sweep.py > ${O}_head_sweep.jsonl 2> ${O}_head_sweep.err; tail -1 ${O}_head_sweep.jsonl | cut -c1-200
timeout 300 python scripts/op_profile.py --batch 256 --top 45 2>&1 | cut -c1-250 > ${O}_op_profile_bf16_b256.txt
timeout 300 python scripts/op_profile.py --batch 256 --precision tf32x3 --top 45 2>&1 | cut -c1-250 > ${O}_op_profile_tf32x3_b256.txt
echo "=== ncu launch lists (one warm step each)"
MTB_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1250 -c 420 --csv --log-file ${O}_launches_bf16.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-frames --no-parity --no-parity-line > ${O}_under_ncu_bf16.log 2>&1; wc -l ${O}_launches_bf16.csv
MTB_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1100 -c 420 --csv --log-file ${O}_launches_tf32x3.csv \
  python bench.py --precision tf32x3 --steps 2 --warmup 3 --no-cpu-baseline --no-frames --no-parity > ${O}_under_ncu_tf32x3.log 2>&1; wc -l ${O}_launches_tf32x3.csv
echo "=== ncu full: tc_conv_kernel (expand, pair projection, stage-1 conv), fmb_kernel, tc32_conv_kernel, depthwise"
timeout 600 ncu --set full --clock-control none -k regex:tc_conv_kernel -c 3 -f -o /tmp/tc_conv python scripts/ncu_ops.py --precision bf16 --batch 256 --ops backbone.1.5.1.block.0,backbone.1.5.1.block.3,backbone.1.1.1.block.0 > ${O}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:fmb_kernel -c 2 -f -o /tmp/fmb python scripts/ncu_ops.py --precision bf16 --batch 256 --fused --ops backbone.1.2.1.block.0,backbone.1.3.1.block.0 > ${O}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:tc32_conv_kernel -c 3 -f -o /tmp/tc32_conv python scripts/ncu_ops.py --precision tf32x3 --batch 128 --ops backbone.1.5.1.block.3,backbone.1.5.1.block.0,backbone.1.2.1.block.0 > ${O}_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:dw3x3s1_tma_kernel -c 1 -f -o /tmp/dw python scripts/ncu_ops.py --precision bf16 --batch 256 --ops backbone.1.5.1.block.1 > ${O}_ncu4.log 2>&1
for r in tc_conv fmb tc32_conv dw; do python scripts/ncu_summary.py /tmp/$r.ncu-rep ${O}_$r.summary.csv; done
ls -la gpurun_out/ | grep r2_final | awk '{print $5, $9}'
