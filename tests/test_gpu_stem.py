"""GPU: the specialised EfficientNet stem kernel (stem3x3s2_kernel: 3x3 / stride 2 / 3 input channels, two output channels
per FFMA2) against conv2d arithmetic (oracle/port_ops.py, /root/reference/metrabs_pytorch/backbones/efficientnet.py:290-293 with
PreprocLayer :1181-1186 folded in) and, bit for bit, against the generic stem kernel it replaces (MTB_STEM_FAST=0 in a
second process: the switch is read once per process)."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import port, port_ops

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_DUMP = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle import port
from tests import helpers
name, side, prec, out = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
pcfg = port.PathConfig(proc_side=side)
sd = port.make_effnet_state_dict(port.effnet_spec(name), pcfg, 8, seed=0, calib_batch=1)
eng = helpers.device_model(name, pcfg, 8, sd, precision=prec).engine()
crops, _ = port.synthetic_inputs(3, side, seed=1)
torch.save(eng.debug_run_ops(crops.cuda(), 1).cpu(), out)
''' % ROOT


@pytest.mark.parametrize('name,side,precision', [('efficientnetv2-l', 256, 'fp32'), ('efficientnetv2-s', 192, 'bf16')])
def test_fast_stem_is_bit_equal_to_generic_stem_and_matches_conv2d(tmp_path, name, side, precision):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    outs = {}
    for fast in ('1', '0'):
        path = str(tmp_path / f'stem_{fast}.pt')
        env = dict(os.environ, MTB_STEM_FAST=fast)
        subprocess.run([sys.executable, '-c', _DUMP, name, str(side), precision, path], check=True, env=env, cwd=ROOT)
        outs[fast] = torch.load(path)
    assert torch.equal(outs['1'], outs['0']), float((outs['1'] - outs['0']).abs().max())
    pcfg = port.PathConfig(proc_side=side)
    spec = port.effnet_spec(name)
    sd = port.make_effnet_state_dict(spec, pcfg, 8, seed=0, calib_batch=1)
    crops, _ = port.synthetic_inputs(3, side, seed=1)
    ref = port_ops.conv_layer_reference(sd, spec, 'backbone.1.0', crops, precision='exact', dtype=torch.float64)
    err = port.relative_error(outs['1'], ref.float())
    print(f'{name}@{side} [{precision}] stem vs conv2d: {err:.2e}')
    assert err < (1e-5 if precision == 'fp32' else 1e-2)
