# epilogue single-slab ring extension + SE cluster kernel variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
prof() { echo "=== $*"; env "$@" timeout 300 python scripts/op_profile.py --batch 256 --top 24 2>&1 | cut -c1-330 | grep "total\|block.3 \|fc1" | cut -c1-400; }
prof A=1
prof MTB_TC_EPI_SINGLE=0
prof MTB_SE_STAGE_W2=1
prof MTB_SE_CB=16
prof MTB_SE_CB=16 MTB_SE_STAGE_W2=1
prof MTB_SE_CLUSTER=0
