#!/bin/bash
# round 2, GPU call 4: ncu --set full of the 3xTF32 kernel on three representative layers (projection / expand / 3x3)
mkdir -p gpurun_out
OPS=backbone.1.5.1.block.3,backbone.1.5.1.block.0,backbone.1.2.1.block.0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc32_conv_kernel -c 3 -f -o gpurun_out/r2_04_tc32 \
   python scripts/ncu_ops.py --precision tf32x3 --batch 128 --ops $OPS > gpurun_out/r2_04_ncu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multiperson.py tests/test_gpu_parity.py -q -s -k "crop_generation or checkpoint" > gpurun_out/r2_04_tests.log 2>&1
tail -3 gpurun_out/r2_04_ncu.log gpurun_out/r2_04_tests.log
