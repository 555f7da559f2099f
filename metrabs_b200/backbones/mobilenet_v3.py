"""MobileNetV3-Small parameter holder for the B200 engine.

Keras-only in the reference (/root/reference/metrabs_tf/backbones/mobilenet_v3.py:258-296, :348-384, :465-553); key schema
defined by this build from the Keras layer names with '/' -> '.': ``backbone.Conv.weight``,
``backbone.Conv.BatchNorm.*``, ``backbone.expanded_conv_<i>.{expand,depthwise,project}.weight`` (+ ``.BatchNorm.*``),
``backbone.expanded_conv_<i>.squeeze_excite.{Conv,Conv_1}.{weight,bias}``, ``backbone.Conv_1.*``, ``backbone.Conv_2.*``."""
from torch import nn

from metrabs_b200 import _lib

_ROWS = [  # (expanded channels, filters, kernel, has SE)
    (16, 16, 3, True), (72, 24, 3, False), (88, 24, 3, False), (96, 40, 5, True), (240, 40, 5, True), (240, 40, 5, True),
    (120, 48, 5, True), (144, 48, 5, True), (288, 96, 5, True), (576, 96, 5, True), (576, 96, 5, True)]


def _depth(v, divisor=8):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def _conv(cin, cout, k, groups=1, bn=True, bias=False):
    m = nn.Conv2d(cin, cout, k, groups=groups, bias=bias)
    if bn:
        m.add_module('BatchNorm', nn.BatchNorm2d(cout, eps=1e-3))
    return m


class Features(nn.Module):
    arch = _lib.ARCH_MOBILENETV3_SMALL
    last_channel = 1024
    stages = []

    def __init__(self):
        super().__init__()
        self.add_module('Conv', _conv(3, 16, 3))
        cin = 16
        for i, (cexp, filters, k, se) in enumerate(_ROWS):
            blk = nn.Module()
            if i != 0:
                blk.add_module('expand', _conv(cin, cexp, 1))
            blk.add_module('depthwise', _conv(cexp, cexp, k, groups=cexp))
            if se:
                sem = nn.Module()
                sem.add_module('Conv', _conv(cexp, _depth(cexp * 0.25), 1, bn=False, bias=True))
                sem.add_module('Conv_1', _conv(_depth(cexp * 0.25), cexp, 1, bn=False, bias=True))
                blk.add_module('squeeze_excite', sem)
            blk.add_module('project', _conv(cexp, filters, 1))
            self.add_module('expanded_conv' if i == 0 else f'expanded_conv_{i}', blk)
            cin = filters
        self.add_module('Conv_1', _conv(cin, _depth(cin * 6), 1))
        self.add_module('Conv_2', _conv(_depth(cin * 6), 1024, 1, bn=False, bias=True))

    def forward(self, x):
        raise RuntimeError('metrabs_b200 backbones run inside Metrabs.forward (libmetrabs_b200.so)')


def mobilenet_v3_small(**kwargs):
    """Use as ``Metrabs(mobilenet_v3_small(), joint_info)``."""
    return Features()
