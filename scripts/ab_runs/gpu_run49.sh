# bench with the CUDA-graph timed region (default) vs plain launches
echo "=== bench default (graph)"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_graph.json | cut -c1-2400
echo "=== bench --graph 0"; timeout 600 python bench.py --no-cpu-baseline --graph 0 2>&1 | tail -1 | cut -c1-200
