#!/bin/bash
# round 2, GPU call 33: fmb pair kernel with an odd number of tiles (dummy tile in the last pair)
timeout 200 python -m pytest tests/test_gpu_fmb.py -x -q -s -k "l-32" 2>&1 | grep -E "fused vs|passed|failed|timed out|Error" | cut -c1-220
