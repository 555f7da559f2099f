"""TEST INFRASTRUCTURE (like oracle/port.py; never imported by metrabs_b200/).

bf16-storage restatement of the EfficientNetV2 crop-model path: the SAME arithmetic as oracle/port.py (which is pinned to
the unmodified reference by tests/golden/), with the roundings of the device's bf16 tensor-core mode inserted at the
points where libmetrabs_b200 rounds:

  * GEMM-type conv weights (1x1 / 3x3, Cin % 8 == 0, Cout % 8 == 0): batch norm folded in fp64, then rounded to bf16
    (csrc/engine.cu prepare_op_weights); stem, depthwise and SE-FC weights stay fp32; biases stay fp32;
  * every activation tensor is stored as bf16: conv / depthwise outputs after bias + activation (+ residual), the SE-scaled
    tensor ahead of the projection (se_scale_kernel), the features; accumulation, SE FCs, logits and decode are fp32;
  * the SE squeeze averages the depthwise outputs BEFORE they are rounded (the depthwise kernel sums its fp32 values).

Purpose: on untrained weights the bf16 mode deviates from the fp32 reference by tens of percent (chaotic amplification of
rounding noise by a non-contractive 170-340-conv network).  This restatement shows that the deviation is what bf16 STORAGE
does to the reference arithmetic - on the CPU, with no kernel of this repo involved - so the device's deviation can be
compared with it instead of with zero (tests/test_oracle_bf16.py, tests/test_gpu_tc.py)."""
import torch
import torch.nn.functional as F

from oracle import port


def _q(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _fold(sd, key, eps, round_w):
    """conv `key`.0 + batch norm `key`.1 -> (weight [Cout,Cin/g,k,k] fp32, bias [Cout] fp32), folded in fp64."""
    w = sd[key + '.0.weight'].double()
    g, b = sd[key + '.1.weight'].double(), sd[key + '.1.bias'].double()
    m, v = sd[key + '.1.running_mean'].double(), sd[key + '.1.running_var'].double()
    s = g / torch.sqrt(v + eps)
    wf = (w * s[:, None, None, None]).float()
    bf = (b - m * s).float()
    return (_q(wf) if round_w else wf), bf


def _tc_like(w, groups, stride):
    cout, cin_g, k, _ = w.shape
    return groups == 1 and k in (1, 3) and stride in (1, 2) and cin_g % 8 == 0 and cout % 8 == 0


def _conv(sd, key, x, stride=1, groups=1, act=True, eps=port.BN_EPS_EFFNETV2):
    """fp32 result of conv + folded BN (+ SiLU), NOT yet rounded (the caller adds the residual first, as the epilogue does)."""
    w, b = _fold(sd, key, eps, _tc_like(sd[key + '.0.weight'], groups, stride))
    y = F.conv2d(x, w, b, stride=stride, groups=groups)
    return F.silu(y) if act else y


def effnet_features_bf16(sd, spec, image, prefix='backbone.1'):
    x = image * 2 - 1
    x = _q(_conv(sd, f'{prefix}.0', port._fixed_pad(x, 3, 0), stride=2))
    for b in port.effnet_block_list(spec):
        key = f'{prefix}.{b["key"]}.block'
        inp = x
        cexp = b['cin'] * b['expand']
        if b['block'] == 'fused':
            x = port._fixed_pad(x, b['kernel'], b['shift'])
            if b['expand'] != 1:
                x = _q(_conv(sd, f'{key}.0', x, stride=b['stride']))
                y = _conv(sd, f'{key}.1', x, act=False)
            else:
                y = _conv(sd, f'{key}.0', x, stride=b['stride'])
        else:
            i = 0
            if b['expand'] != 1:
                x = _q(_conv(sd, f'{key}.{i}', x))
                i += 1
            x = port._fixed_pad(x, b['kernel'], b['shift'])
            d = _conv(sd, f'{key}.{i}', x, stride=b['stride'], groups=cexp)  # fp32 depthwise output
            i += 1
            s = d.mean(dim=(2, 3), keepdim=True)                               # squeeze: before the bf16 rounding
            s = F.silu(F.conv2d(s, sd[f'{key}.{i}.fc1.weight'], sd[f'{key}.{i}.fc1.bias']))
            s = torch.sigmoid(F.conv2d(s, sd[f'{key}.{i}.fc2.weight'], sd[f'{key}.{i}.fc2.bias']))
            x = _q(_q(d) * s)                                                  # stored bf16, then se_scale_kernel rounds again
            i += 1
            y = _conv(sd, f'{key}.{i}', x, act=False)
        x = _q(y + inp) if b['residual'] else _q(y)
    n_stage = len(spec.stages)
    return _q(_conv(sd, f'{prefix}.{n_stage + 1}', x))


def metrabs_forward_bf16(sd, spec, cfg, n_joints, image, intrinsics, stages=None):
    """oracle/port.metrabs_forward with bf16 storage emulated (features bf16, head weights bf16, decode fp32)."""
    features = effnet_features_bf16(sd, spec, image)
    sd_h = dict(sd)
    sd_h['heatmap_heads.conv_final.weight'] = _q(sd['heatmap_heads.conv_final.weight'])
    coords2d, coords3d_rel = port.heads(sd_h, features, cfg, n_joints)
    out = port.reconstruct_absolute(coords2d, coords3d_rel, intrinsics, cfg)
    if stages is not None:
        stages.update(features=features, coords2d=coords2d, coords3d_rel=coords3d_rel)
    return out
