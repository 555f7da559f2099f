"""TEST INFRASTRUCTURE ONLY - torch-cpu restatement of the two backbones that exist ONLY as TF/Keras code in the reference:

* ResNet-50 V1 with the MeTRAbs train/test stride-dilation switching   /root/reference/metrabs_tf/backbones/resnet.py
    stem + pool :170-198, bottleneck ``block1_dense`` :239-319 (V1: stride on the first 1x1 and on the shortcut 1x1, bias
    on every conv, BN eps 1e-5 :71), ``stack1_dense`` :515-537, stride plan ``get_strides_and_dilations`` :601-618 and
    ``ResNetUnified`` :621-666 (V1 uses ``dil_out`` for the first block of a stack too, :636-644),
    preprocessing ``caffe_preproc`` backbones/builder.py:106-108.
* MobileNetV3-Small  /root/reference/metrabs_tf/backbones/mobilenet_v3.py  stem / Conv_1 / Conv_2 :258-296, table
    :364-384, ``_inverted_res_block`` :490-553, ``_se_block`` :465-487, ``correct_pad`` :556-575, ``_depth`` :449-456,
    preprocessing builder.py:116-117 (x*255) followed by the in-model ``Rescaling(1/127.5, -1)`` :259.

PARITY UNPINNED: the reference holds no tests/goldens for these, TensorFlow/Keras are not installed, and part of the
arithmetic lives in the un-vendored ``fleras`` package (``Conv2DDenseSame``, pinned only as
``git+https://github.com/isarandi/fleras.git``, environment.yml:43).  ``Conv2DDenseSame(strides, bottomright_stride)`` is
restated as "SAME-padded conv evaluated at pixels ``shift::stride``" (shift = 1 for bottomright), which is how the
vendored EfficientNet code realises the same flag (effnetv2_utils.py:122-129).  Device-vs-oracle parity for these two
backbones is therefore "build's restatement vs build's kernels".

Key schema (this build's; Keras layer names, '/' -> '.', torch tensor layouts): ``backbone.<layer>.weight|bias`` for convs,
``backbone.<bn layer>.weight|bias|running_mean|running_var`` for BN.
"""
import math

import torch
import torch.nn.functional as F

from oracle import port

RESNET_BN_EPS = 1e-5
MOBILENET_BN_EPS = 1e-3


def _bn(sd, key, x, eps):
    return F.batch_norm(x, sd[key + '.running_mean'], sd[key + '.running_var'], sd[key + '.weight'], sd[key + '.bias'],
                        training=False, eps=eps)


def hard_sigmoid(x):
    return F.relu6(x + 3.0) * (1.0 / 6.0)


def hard_swish(x):
    return hard_sigmoid(x) * x


# ----------------------------------------------------------------------------------------------------- ResNet-50
def resnet_stride_plan(output_stride, centered_stride):
    """get_strides_and_dilations (resnet.py:601-618)."""
    brs = [False, False, False]
    i_last = int(round(math.log2(output_stride))) - 3
    if centered_stride and i_last >= 0:
        brs[i_last] = True
    dil_in, dil_out, strides = [1, 1, 1], [1, 1, 1], [2, 2, 2]
    for i in range(max(0, i_last + 1), 3):
        strides[i] = 1
        dil_in[i] = 2 ** (i - (i_last + 1))
        dil_out[i] = dil_in[i] * 2
    return strides, dil_in, dil_out, brs


def resnet50_blocks(cfg: port.PathConfig):
    """[(name, filters, stride, shift, dilation, conv_shortcut)] in execution order (inference: stride_test)."""
    strides, dil_in, dil_out, brs = resnet_stride_plan(cfg.stride_test, cfg.centered_stride)
    out = []
    for st, (f, n) in enumerate(zip([64, 128, 256, 512], [3, 4, 6, 3])):
        for bi in range(n):
            first = bi == 0
            stride = strides[st - 1] if (st > 0 and first) else 1
            shift = 1 if (st > 0 and first and brs[st - 1]) else 0
            dil = dil_in[0] if st == 0 else dil_out[st - 1]
            out.append((f'conv{st + 2}_block{bi + 1}', f, stride, shift, dil, first))
    return out


class ResNet50Spec:
    name = 'resnet50'
    out_channels = 2048

    def __init__(self, cfg: port.PathConfig):
        self.cfg = cfg

    def features(self, sd, image, tap=None, init=None):
        """[B,3,S,S] in [0,1] -> [B,2048,S/s,S/s].  With ``init`` = (generator) the weights are created and BN-calibrated
        on the fly (conditioned random init), otherwise read from ``sd``."""
        p = 'backbone.'
        g = init

        def conv_bn(x, cname, bname, cout, k, stride=1, shift=0, dil=1, pad=0, relu=True, damp=1.0):
            if g is not None:
                cin = x.shape[1]
                sd[p + cname + '.weight'] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))
                sd[p + cname + '.bias'] = 0.1 * torch.randn(cout, generator=g)
            y = F.conv2d(x, sd[p + cname + '.weight'], sd[p + cname + '.bias'], padding=pad, dilation=dil)
            if stride > 1 or shift:
                y = y[:, :, shift::stride, shift::stride]  # Conv2DDenseSame: dense SAME conv sampled at shift::stride
            if g is not None:
                port._calibrate_bn(sd, p + bname, y, g, RESNET_BN_EPS, damp)
            y = _bn(sd, p + bname, y, RESNET_BN_EPS)
            y = F.relu(y) if relu else y
            if tap is not None:
                tap[p + cname] = y
            return y

        mean = torch.tensor([103.939, 116.779, 123.68]).reshape(1, 3, 1, 1)
        x = 255.0 * image - mean  # caffe_preproc, no channel swap
        x = conv_bn(F.pad(x, (3, 3, 3, 3)), 'conv1_conv', 'conv1_bn', 64, 7, stride=2)
        x = F.max_pool2d(F.pad(x, (1, 1, 1, 1)), 3, stride=2)  # zero pad (post-ReLU values are >= 0), then VALID
        if tap is not None:
            tap[p + 'pool1_pool'] = x
        for name, f, stride, shift, dil, conv_shortcut in resnet50_blocks(self.cfg):
            inp = x
            sc = conv_bn(inp, name + '_0_conv', name + '_0_bn', 4 * f, 1, stride, shift, relu=False) if conv_shortcut else inp
            y = conv_bn(inp, name + '_1_conv', name + '_1_bn', f, 1, stride, shift)
            y = conv_bn(y, name + '_2_conv', name + '_2_bn', f, 3, dil=dil, pad=dil)
            y = conv_bn(y, name + '_3_conv', name + '_3_bn', 4 * f, 1, relu=False, damp=0.5)
            x = F.relu(sc + y)
            if tap is not None:
                tap[p + name + '_3_conv'] = x
        return x


# ------------------------------------------------------------------------------------------------ MobileNetV3-Small
def _depth(v, divisor=8):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


MOBILENETV3_SMALL_ROWS = [
    # (expansion, filters, kernel, stride, se, activation, bottomright)
    (1, 16, 3, 2, True, 'relu', False), (72. / 16, 24, 3, 2, False, 'relu', False), (88. / 24, 24, 3, 1, False, 'relu', False),
    (4, 40, 5, 2, True, 'hswish', False), (6, 40, 5, 1, True, 'hswish', False), (6, 40, 5, 1, True, 'hswish', False),
    (3, 48, 5, 1, True, 'hswish', False), (3, 48, 5, 1, True, 'hswish', False), (6, 96, 5, 2, True, 'hswish', True),
    (6, 96, 5, 1, True, 'hswish', False), (6, 96, 5, 1, True, 'hswish', False)]


class MobileNetV3SmallSpec:
    name = 'mobilenetv3-small'
    out_channels = 1024

    def __init__(self, cfg: port.PathConfig):
        self.cfg = cfg

    def features(self, sd, image, tap=None, init=None):
        p = 'backbone.'
        g = init
        acts = {'relu': F.relu, 'hswish': hard_swish}

        def conv_bn(x, cname, cout, k, stride=1, groups=1, act=None, bn=True, bias=False, damp=1.0):
            if g is not None:
                cin = x.shape[1] // groups
                sd[p + cname + '.weight'] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))
                if bias:
                    sd[p + cname + '.bias'] = 0.1 * torch.randn(cout, generator=g)
            y = F.conv2d(x, sd[p + cname + '.weight'], sd[p + cname + '.bias'] if bias else None, stride=stride, groups=groups)
            if bn:
                if g is not None:
                    port._calibrate_bn(sd, p + cname + '.BatchNorm', y, g, MOBILENET_BN_EPS, damp)
                y = _bn(sd, p + cname + '.BatchNorm', y, MOBILENET_BN_EPS)
            y = act(y) if act is not None else y
            if tap is not None:
                tap[p + cname] = y
            return y

        x = image * 2 - 1  # 255*x (builder.py:116-117) then Rescaling(1/127.5, -1) (mobilenet_v3.py:259)
        s = x.shape[-1]
        out = (s + 1) // 2
        pad_total = max((out - 1) * 2 + 3 - s, 0)  # TF 'same', stride 2
        pb = pad_total // 2
        x = conv_bn(F.pad(x, (pb, pad_total - pb, pb, pad_total - pb)), 'Conv', 16, 3, stride=2, act=hard_swish)
        for bi, (exp, filters, k, stride, se, act, br) in enumerate(MOBILENETV3_SMALL_ROWS):
            name = 'expanded_conv' if bi == 0 else f'expanded_conv_{bi}'
            a = acts[act]
            inp = x
            cin = x.shape[1]
            cexp = _depth(cin * exp)
            if bi != 0:
                x = conv_bn(x, name + '.expand', cexp, 1, act=a)
            shift = 1 if (br and self.cfg.centered_stride) else 0
            pbeg = (k - 1) // 2
            pend = k - 1 - pbeg
            if stride == 2:
                x = F.pad(x, (pbeg - shift, pend + shift, pbeg - shift, pend + shift))  # correct_pad, then VALID
            else:
                x = F.pad(x, (pbeg, pend, pbeg, pend))  # 'same', stride 1
            x = conv_bn(x, name + '.depthwise', cexp, k, stride=stride, groups=cexp, act=a)
            if se:
                csq = _depth(cexp * 0.25)
                if g is not None:
                    sd[p + name + '.squeeze_excite.Conv.weight'] = torch.randn(csq, cexp, 1, 1, generator=g) * math.sqrt(2.0 / cexp)
                    sd[p + name + '.squeeze_excite.Conv.bias'] = 0.2 * torch.randn(csq, generator=g)
                    sd[p + name + '.squeeze_excite.Conv_1.weight'] = torch.randn(cexp, csq, 1, 1, generator=g) * math.sqrt(2.0 / csq)
                    sd[p + name + '.squeeze_excite.Conv_1.bias'] = 1.0 * torch.randn(cexp, generator=g)
                q = x.mean(dim=(2, 3), keepdim=True)
                q = F.relu(F.conv2d(q, sd[p + name + '.squeeze_excite.Conv.weight'], sd[p + name + '.squeeze_excite.Conv.bias']))
                q = hard_sigmoid(F.conv2d(q, sd[p + name + '.squeeze_excite.Conv_1.weight'],
                                          sd[p + name + '.squeeze_excite.Conv_1.bias']))
                x = x * q
            res = stride == 1 and cin == filters
            x = conv_bn(x, name + '.project', filters, 1, damp=0.5 if res else 1.0)
            if res:
                x = x + inp
                if tap is not None:
                    tap[p + name + '.project'] = x
        x = conv_bn(x, 'Conv_1', _depth(x.shape[1] * 6), 1, act=hard_swish)
        x = conv_bn(x, 'Conv_2', 1024, 1, act=hard_swish, bn=False, bias=True)
        return x


def make_state_dict(spec, cfg: port.PathConfig, n_joints, seed=0, calib_batch=4, head_gain=10.0):
    """Conditioned random init for the TF-only backbones (same recipe as port.make_effnet_state_dict)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    calib, _ = port.synthetic_inputs(calib_batch, cfg.proc_side, seed=seed + 77)
    with torch.no_grad():
        spec.features(sd, calib, init=g)
    port.init_head(sd, g, spec.out_channels, n_joints, cfg.depth, head_gain)
    return sd
