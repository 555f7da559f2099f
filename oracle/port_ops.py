"""TEST INFRASTRUCTURE ONLY - per-layer reference arithmetic for the device kernels' unit tests.

``oracle/port.py`` restates the whole path; this file exposes ONE conv layer of it at a time (same reference lines:
``/root/reference/metrabs_pytorch/backbones/efficientnet.py`` :110-173 MBConv, :176-234 FusedMBConv, :290-293 stem,
:319-324 last conv, :1127-1161 fixed padding; eval-mode BatchNorm eps 1e-3 :1051) so that a single device launch
(``mtb_debug_run_op``) can be compared with plain ``torch.nn.functional.conv2d`` arithmetic on identical operands instead
of with another kernel of this repository.

``precision``:
* ``'exact'``  conv -> BN (eval) -> act -> (+res) evaluated in the requested dtype (fp64 on the CPU, fp32 on the GPU with
               TF32 disabled): the bar for the fp32 / 3xTF32 kernels.
* ``'bf16'``   the SAME arithmetic at the roundings the bf16 tensor-core mode defines: BN folded into the conv weight in
               fp64 and rounded ONCE to bf16 (``w*gamma/sqrt(var+eps)``), bf16 input (and bf16(scale*x) for a
               squeeze-excitation projection), wide accumulation, fp32 bias, act, +res; the caller rounds the result to
               bf16 or allows one bf16 ulp.
"""
import torch
import torch.nn.functional as F

from oracle import port


def effnet_op_table(spec: port.EffNetSpec, prefix='backbone.1'):
    """engine op name (= reference key prefix of the layer) -> dict(stride, shift, act, depthwise, kernel)."""
    t = {f'{prefix}.0': dict(stride=2, shift=0, act=True, depthwise=False, kernel=3, stem=True)}
    for b in port.effnet_block_list(spec):
        key = f'{prefix}.{b["key"]}.block'
        if b['block'] == 'fused':
            t[f'{key}.0'] = dict(stride=b['stride'], shift=b['shift'], act=True, depthwise=False, kernel=b['kernel'])
            if b['expand'] != 1:
                t[f'{key}.1'] = dict(stride=1, shift=0, act=False, depthwise=False, kernel=1)
        else:
            i = 0
            if b['expand'] != 1:
                t[f'{key}.{i}'] = dict(stride=1, shift=0, act=True, depthwise=False, kernel=1)
                i += 1
            t[f'{key}.{i}'] = dict(stride=b['stride'], shift=b['shift'], act=True, depthwise=True, kernel=b['kernel'])
            i += 2  # squeeze-excitation sits between the depthwise conv and the projection
            t[f'{key}.{i}'] = dict(stride=1, shift=0, act=False, depthwise=False, kernel=1)
    t[f'{prefix}.{len(spec.stages) + 1}'] = dict(stride=1, shift=0, act=True, depthwise=False, kernel=1)
    return t


def fold_conv_bn(sd, key, eps=port.BN_EPS_EFFNETV2):
    """Conv2dNormActivation in eval mode as one affine conv: (w * g/sqrt(v+eps), b - m*g/sqrt(v+eps)), in fp64."""
    w = sd[key + '.0.weight'].double()
    g, b = sd[key + '.1.weight'].double(), sd[key + '.1.bias'].double()
    m, v = sd[key + '.1.running_mean'].double(), sd[key + '.1.running_var'].double()
    s = g / torch.sqrt(v + eps)
    return w * s[:, None, None, None], b - m * s


def conv_layer_reference(sd, spec, name, x_nhwc, res_nhwc=None, scale=None, precision='exact', dtype=torch.float64):
    """One conv layer of EfficientNet.features on ``x_nhwc`` [B,H,W,C] (the stem takes NCHW crops in [0,1] and applies
    PreprocLayer x*2-1, efficientnet.py:1181-1186).  Returns NHWC in ``dtype``."""
    op = effnet_op_table(spec)[name]
    dev = x_nhwc.device
    w, bias = fold_conv_bn(sd, name)
    if precision == 'bf16' and not op['depthwise'] and not op.get('stem'):
        w = w.float().bfloat16().double()  # the tensor-core weights; depthwise / stem weights stay fp32 on the device
    w, bias = w.to(dev, dtype), bias.to(dev, dtype)
    if op.get('stem'):
        x = x_nhwc.to(dtype) * 2 - 1
    else:
        x = x_nhwc.permute(0, 3, 1, 2).to(dtype)
    if scale is not None:
        x = x * scale.to(dev, dtype)[:, :, None, None]
        if precision == 'bf16':
            x = x.float().bfloat16().to(dtype)
    if op['kernel'] > 1:
        x = port._fixed_pad(x, op['kernel'], op['shift'])
    y = F.conv2d(x, w, bias, stride=op['stride'], groups=x.shape[1] if op['depthwise'] else 1)
    if op['act']:
        y = F.silu(y)
    y = y.permute(0, 2, 3, 1)
    if res_nhwc is not None:
        y = y + res_nhwc.to(dtype)
    return y.contiguous()
