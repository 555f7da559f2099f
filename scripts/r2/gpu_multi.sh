#!/bin/bash
# round-2 multi-GPU run (gpurun --gpus N): NCCL test of the sharded forward, weak-scaling and strong-scaling (config c3) bench lines
mkdir -p gpurun_out
N=${1:-2}
O=gpurun_out/r2_multi_n$N
timeout 900 python -m pytest tests/test_gpu_multi.py -q -s > ${O}_tests.log 2>&1; tail -2 ${O}_tests.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29611 --steps 10 --warmup 3 2> ${O}_bench.err | tail -1 > ${O}_bench_weak.json; cut -c1-260 ${O}_bench_weak.json
run 29612 --steps 10 --warmup 3 --scaling strong --side 384 --batch 256 --no-parity-line 2>> ${O}_bench.err | tail -1 > ${O}_bench_strong_c3.json; cut -c1-260 ${O}_bench_strong_c3.json
timeout 600 python bench.py --side 384 --batch 256 --steps 10 --warmup 3 --no-cpu-baseline --no-frames --no-parity-line 2>> ${O}_bench.err | tail -1 > ${O}_bench_c3_n1.json; cut -c1-200 ${O}_bench_c3_n1.json
