#!/bin/bash
# round 2, GPU call 28: fmb_kernel on the tiny model (Cin = 16, K1 = 1, single 64-channel chunk) + smoke with / without fusion
mkdir -p gpurun_out
O=gpurun_out/r2_28
timeout 300 python -m pytest tests/test_gpu_fmb.py -x -q -s -k "tiny" > ${O}_tests.log 2>&1; grep -E "fused vs|passed|failed|assert|Error" ${O}_tests.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
MTB_FMB=0 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke | sed 's/^/MTB_FMB=0 /'
