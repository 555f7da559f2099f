#!/bin/bash
# round-2 8-GPU run: weak-scaling headline and strong-scaling config c3 as written (EfficientNetV2-L@384, 256 crops in total -> 32 per GPU)
mkdir -p gpurun_out
N=8
O=gpurun_out/r2_multi_n$N
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29621 --steps 10 --warmup 3 --no-parity-line --no-frames 2> ${O}_bench.err | tail -1 > ${O}_bench_weak.json; cut -c1-260 ${O}_bench_weak.json
run 29622 --steps 10 --warmup 3 --scaling strong --side 384 --batch 256 --no-parity-line --no-frames 2>> ${O}_bench.err | tail -1 > ${O}_bench_strong_c3.json; cut -c1-260 ${O}_bench_strong_c3.json
