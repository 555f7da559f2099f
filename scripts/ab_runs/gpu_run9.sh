P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $P scripts/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -8
echo "=== bench N=2"; timeout 900 $P bench.py --gpus 2 --steps 10 --warmup 3 --batch 128 2>&1 | grep -v "^W\|^\*\*\*" | cut -c1-900 | tail -4
echo "=== bench N=1 (same box)"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --batch 128 --no-cpu-baseline 2>&1 | cut -c1-300 | tail -3
