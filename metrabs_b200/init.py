"""Calibration-free conditioned random initialisation for synthetic-weight runs (bench.py, demos).

torch's default init leaves an eval-mode EfficientNetV2 with dead features (sigma ~ 1e-7), which makes
every joint decode to the volume centre.  This recipe keeps activations O(1) through the stack without running the
network: fan-in scaled normal conv weights with a SiLU gain, identity BN statistics, damped residual branches and a
head scaled for peaky heatmaps.  (The parity fixtures use the oracle's calibrated recipe instead.)"""
import math

import torch
from torch import nn


@torch.no_grad()
def conditioned_random_init_(model, seed=0, head_gain=8.0):
    g = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, nn.Conv2d):
            fan_in = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3]
            gain = 1.6
            if name.endswith('conv_final'):
                gain = head_gain
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (gain / math.sqrt(fan_in)))
            if m.bias is not None:
                m.bias.copy_((torch.rand(m.bias.shape, generator=g) - 0.5) * 0.1)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
            m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            m.running_mean.zero_()
            m.running_var.fill_(1.0)
    # damp the last BN of every residual branch so the residual sum does not blow up with depth
    for name, m in model.named_modules():
        if name.endswith('.block'):
            last = [c for c in m.children()][-1]
            bns = [c for c in last.modules() if isinstance(c, nn.BatchNorm2d)]
            if bns:
                bns[-1].weight.mul_(0.3)
    if hasattr(model, 'mark_weights_changed'):
        model.mark_weights_changed()
    return model
