"""TEST INFRASTRUCTURE ONLY - CPU restatement ("oracle port") of the MeTRAbs per-crop inference hot path.

This file is the checker for the CUDA path, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  ``metrabs_b200`` never does.

It restates, in plain torch-cpu fp32 functional code over a flat ``state_dict`` in the reference key schema
(``backbone.1.<stage>.<block>.block.<i>...``, ``heatmap_heads.conv_final.*``), what the reference computes in

* ``/root/reference/metrabs_pytorch/backbones/efficientnet.py`` :110-173 (MBConv), :176-234 (FusedMBConv),
  :237-357 (EfficientNet.features), :379-433 (configs), :1127-1161 (fixed padding), :1181-1186 (PreprocLayer)
* ``/root/reference/metrabs_pytorch/models/metrabs.py`` :47-64 (Metrabs.forward), :67-85 (MetrabsHeads)
* ``/root/reference/metrabs_pytorch/models/util.py`` :6-33 (heatmap_to_image / heatmap_to_metric)
* ``/root/reference/metrabs_pytorch/ptu.py`` :47-92 (softmax / soft_argmax / decode_heatmap / linspace)
* ``/root/reference/metrabs_pytorch/ptu3d.py`` :9-33, :52-121 (reconstruct_absolute and helpers)

Parity pin: the reference holds no tests or golden vectors for this path, so the pin is the
reference itself, imported unmodified in the build container by ``oracle/gen_golden.py`` (see
``oracle/ref_import.py``); the tensors it produced are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` checks this port against them.  ResNet-50 / MobileNetV3 exist only as TF/Keras
code that depends on the un-vendored ``fleras`` package, so their restatement (``oracle/port_tf_backbones.py``)
is "parity unpinned" and says so.
"""
import dataclasses
import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# configuration (the keys the path reads from get_config(); reference config/config_l.yaml:1-21)
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class PathConfig:
    proc_side: int = 256
    stride_train: int = 32
    stride_test: int = 32
    centered_stride: bool = True
    legacy_centered_stride_bug: bool = False
    depth: int = 8
    box_size_mm: float = 2200.0
    weak_perspective: bool = False
    mix_3d_inside_fov: Optional[float] = 0.5
    affine_weights: Optional[str] = None
    transform_coords: bool = False
    predict_all_and_latents: bool = False

    def as_reference_dict(self):
        return dataclasses.asdict(self)


@dataclasses.dataclass
class StageSpec:
    """One row of the reference's ``inverted_residual_setting`` (efficientnet.py:379-433)."""
    block: str  # 'fused' or 'mb'
    expand: int
    kernel: int
    stride: int
    cin: int
    cout: int
    layers: int
    bottomright: bool = False


@dataclasses.dataclass
class EffNetSpec:
    name: str
    stages: List[StageSpec]
    last_channel: int

    @property
    def stem_channels(self):
        return self.stages[0].cin


def effnet_spec(name, centered_stride=True):
    """EfficientNetV2 S/M/L tables (efficientnet.py:398-429) plus a 'tiny' table of the same block grammar that
    is small enough to commit full weights as a golden fixture."""
    f, m = 'fused', 'mb'
    br = bool(centered_stride)
    if name == 'efficientnetv2-s':
        rows = [(f, 1, 3, 1, 24, 24, 2), (f, 4, 3, 2, 24, 48, 4), (f, 4, 3, 2, 48, 64, 4),
                (m, 4, 3, 2, 64, 128, 6), (m, 6, 3, 1, 128, 160, 9), (m, 6, 3, 2, 160, 256, 15, br)]
    elif name == 'efficientnetv2-m':
        rows = [(f, 1, 3, 1, 24, 24, 3), (f, 4, 3, 2, 24, 48, 5), (f, 4, 3, 2, 48, 80, 5),
                (m, 4, 3, 2, 80, 160, 7), (m, 6, 3, 1, 160, 176, 14), (m, 6, 3, 2, 176, 304, 18, br),
                (m, 6, 3, 1, 304, 512, 5)]
    elif name == 'efficientnetv2-l':
        rows = [(f, 1, 3, 1, 32, 32, 4), (f, 4, 3, 2, 32, 64, 7), (f, 4, 3, 2, 64, 96, 7),
                (m, 4, 3, 2, 96, 192, 10), (m, 6, 3, 1, 192, 224, 19), (m, 6, 3, 2, 224, 384, 25, br),
                (m, 6, 3, 1, 384, 640, 7)]
    elif name == 'efficientnetv2-tiny':
        rows = [(f, 1, 3, 1, 8, 8, 1), (f, 4, 3, 2, 8, 16, 2), (f, 4, 3, 2, 16, 24, 1),
                (m, 4, 3, 2, 24, 32, 2), (m, 6, 3, 1, 32, 40, 1), (m, 6, 3, 2, 40, 48, 2, br)]
        return EffNetSpec(name, [StageSpec(*r) for r in rows], last_channel=64)
    else:
        raise ValueError(name)
    return EffNetSpec(name, [StageSpec(*r) for r in rows], last_channel=1280)


# ----------------------------------------------------------------------------------------------
# backbone (efficientnet.py)
# ----------------------------------------------------------------------------------------------
BN_EPS_EFFNETV2 = 1e-3  # efficientnet.py:1051 (norm_layer=partial(BatchNorm2d, eps=1e-3))


def _fixed_pad(x, kernel, shift):
    """efficientnet.py:1127-1161: explicit zero pad (pb-shift, pe+shift) on both spatial axes, then VALID."""
    total = kernel - 1
    pb = total // 2
    pe = total - pb
    return F.pad(x, (pb - shift, pe + shift, pb - shift, pe + shift))


def _bn(sd, key, x, eps):
    return F.batch_norm(x, sd[key + '.running_mean'], sd[key + '.running_var'], sd[key + '.weight'],
                        sd[key + '.bias'], training=False, eps=eps)


def _conv_bn(sd, key, x, stride=1, groups=1, act=True, eps=BN_EPS_EFFNETV2, tap=None):
    """Conv2dNormActivation: conv (no bias) -> BN (eval) -> optional SiLU.  ``key``.0 = conv, ``key``.1 = BN."""
    x = F.conv2d(x, sd[key + '.0.weight'], None, stride=stride, groups=groups)
    x = _bn(sd, key + '.1', x, eps)
    if act:
        x = F.silu(x)
    if tap is not None:
        tap[key] = x
    return x


def effnet_block_list(spec: EffNetSpec):
    """Flattened per-block descriptors in execution order, with their reference key prefixes."""
    blocks = []
    for si, st in enumerate(spec.stages):
        for bi in range(st.layers):
            first = bi == 0
            cin = st.cin if first else st.cout
            blocks.append(dict(
                key=f'{si + 1}.{bi}', block=st.block, expand=st.expand, kernel=st.kernel,
                stride=st.stride if first else 1, cin=cin, cout=st.cout,
                shift=1 if (first and st.bottomright) else 0,
                residual=(first and st.stride == 1 and st.cin == st.cout) or (not first)))
    return blocks


def effnet_features(sd, spec: EffNetSpec, image, prefix='backbone.1', tap=None):
    """[B,3,S,S] fp32 in [0,1] -> [B,last_channel,S/32,S/32]  (PreprocLayer + EfficientNet.features)."""
    x = image * 2 - 1  # efficientnet.py:1185
    x = _conv_bn(sd, f'{prefix}.0', _fixed_pad(x, 3, 0), stride=2, tap=tap)  # :290-293
    for b in effnet_block_list(spec):
        key = f'{prefix}.{b["key"]}.block'
        inp = x
        cexp = b['cin'] * b['expand']
        if b['block'] == 'fused':  # :176-234
            x = _fixed_pad(x, b['kernel'], b['shift'])
            if b['expand'] != 1:
                x = _conv_bn(sd, f'{key}.0', x, stride=b['stride'], tap=tap)
                x = _conv_bn(sd, f'{key}.1', x, act=False, tap=tap)
            else:
                x = _conv_bn(sd, f'{key}.0', x, stride=b['stride'], tap=tap)
        else:  # :110-173
            i = 0
            if b['expand'] != 1:
                x = _conv_bn(sd, f'{key}.{i}', x, tap=tap)
                i += 1
            x = _fixed_pad(x, b['kernel'], b['shift'])
            x = _conv_bn(sd, f'{key}.{i}', x, stride=b['stride'], groups=cexp, tap=tap)
            i += 1
            # torchvision SqueezeExcitation: avgpool -> fc1 -> SiLU -> fc2 -> sigmoid -> scale
            s = x.mean(dim=(2, 3), keepdim=True)
            s = F.silu(F.conv2d(s, sd[f'{key}.{i}.fc1.weight'], sd[f'{key}.{i}.fc1.bias']))
            s = torch.sigmoid(F.conv2d(s, sd[f'{key}.{i}.fc2.weight'], sd[f'{key}.{i}.fc2.bias']))
            x = x * s
            i += 1
            x = _conv_bn(sd, f'{key}.{i}', x, act=False, tap=tap)
        if b['residual']:
            x = x + inp  # StochasticDepth is the identity in eval mode
        if tap is not None:
            tap[f'{prefix}.{b["key"]}'] = x
    n_stage = len(spec.stages)
    x = _conv_bn(sd, f'{prefix}.{n_stage + 1}', x, tap=tap)  # :319-324
    return x


# ----------------------------------------------------------------------------------------------
# head + decode (models/metrabs.py:67-85, ptu.py:47-92, models/util.py:6-33)
# ----------------------------------------------------------------------------------------------
def linspace01(n, dtype=torch.float32):
    """ptu.py:78-92 with start=0, stop=1, endpoint=True: num==1 -> [0.5]."""
    if n == 1:
        return torch.full((1,), 0.5, dtype=dtype)
    return torch.linspace(0.0, 1.0, n, dtype=dtype)


def soft_argmax(logits, dims):
    """Joint softmax over ``dims`` then per-axis expectation with linspace(0,1,n); output coordinate order follows
    ``dims`` (ptu.py:47-75).  Returns [..., len(dims)] with the heatmap axes removed."""
    dims = tuple(d if d >= 0 else logits.ndim + d for d in dims)
    mx = torch.amax(logits, dim=dims, keepdim=True)
    e = torch.exp(logits - mx)
    p = e / torch.sum(e, dim=dims, keepdim=True)
    out = []
    for d in dims:
        others = [o for o in dims if o != d]
        marg = torch.sum(p, dim=others, keepdim=True) if others else p
        shape = [1] * logits.ndim
        shape[d] = logits.shape[d]
        coord = (marg * linspace01(logits.shape[d], logits.dtype).reshape(shape)).sum(dim=d, keepdim=True)
        for hd in sorted(dims, reverse=True):
            coord = coord.squeeze(hd)
        out.append(coord)
    return torch.stack(out, dim=-1)


def heatmap_to_image(coords, cfg: PathConfig, is_training=False):
    """models/util.py:6-20."""
    stride = cfg.stride_train if is_training else cfg.stride_test
    last_image_pixel = cfg.proc_side - 1
    last_receptive_center = last_image_pixel - (last_image_pixel % stride)
    out = coords * last_receptive_center
    if cfg.centered_stride:
        out = out + stride // 2
    if cfg.legacy_centered_stride_bug:
        out = out + stride // 2
    return out


def heatmap_to_metric(coords, cfg: PathConfig, is_training=False):
    """models/util.py:29-33."""
    xy = heatmap_to_image(coords[..., :2], cfg, is_training) * cfg.box_size_mm / cfg.proc_side
    return torch.cat([xy, coords[..., 2:] * cfg.box_size_mm], dim=-1)


def head_logits(sd, features, prefix='heatmap_heads.conv_final'):
    """1x1 conv with bias, [B,C,H,W] -> [B,J+D*J,H,W] (models/metrabs.py:73,76)."""
    return F.conv2d(features, sd[prefix + '.weight'], sd[prefix + '.bias'])


def split_logits(x, n_joints, depth):
    """models/metrabs.py:78-79: channel = J + d*J + j."""
    b, _, h, w = x.shape
    logits2d = x[:, :n_joints]
    logits3d = x[:, n_joints:].reshape(b, depth, n_joints, h, w)
    return logits2d, logits3d


def heads(sd, features, cfg: PathConfig, n_joints):
    """MetrabsHeads.forward (models/metrabs.py:75-85) -> (coords2d [B,J,2] px, coords3d_rel [B,J,3] mm)."""
    x = head_logits(sd, features)
    logits2d, logits3d = split_logits(x, n_joints, cfg.depth)
    coords3d = soft_argmax(logits3d.float(), dims=(4, 3, 1))
    coords3d_rel = heatmap_to_metric(coords3d, cfg)
    coords2d = soft_argmax(logits2d.float(), dims=(3, 2))
    coords2d_px = heatmap_to_image(coords2d, cfg)
    return coords2d_px, coords3d_rel


# ----------------------------------------------------------------------------------------------
# absolute reconstruction (ptu3d.py)
# ----------------------------------------------------------------------------------------------
def to_homogeneous(x):
    return torch.cat([x, torch.ones_like(x[..., :1])], dim=-1)


def is_within_fov(imcoords, cfg: PathConfig, border_factor=0.75):
    """ptu3d.py:113-121 (bounds inclusive)."""
    offset = -cfg.stride_train / 2 if not cfg.centered_stride else 0
    lower = cfg.stride_train * border_factor + offset
    upper = cfg.proc_side - cfg.stride_train * border_factor + offset
    return torch.all(torch.logical_and(imcoords >= lower, imcoords <= upper), dim=-1)


def rms_scale(x):
    """ptu3d.py:71-74: RMS over the WHOLE tensor (i.e. over the batch too)."""
    return x.square().mean().sqrt()


def reconstruct_ref_fullpersp(normalized_2d, coords3d_rel, validity_mask):
    """ptu3d.py:56-105: weighted ridge least squares for the reference point, rows scaled by batch-global RMS."""
    nb, nj = normalized_2d.shape[:2]
    scale2d = rms_scale(normalized_2d)
    x2 = (normalized_2d / scale2d).reshape(nb, nj * 2, 1)
    eyes2 = torch.eye(2, dtype=normalized_2d.dtype).repeat(nb, nj, 1)
    a = torch.cat([eyes2, -x2], dim=2)
    a = torch.cat([a, torch.eye(3, dtype=a.dtype).expand(nb, 3, 3)], dim=1)
    rel_backproj = normalized_2d * coords3d_rel[:, :, 2:] - coords3d_rel[:, :, :2]
    scale_b = rms_scale(rel_backproj)
    b = (rel_backproj / scale_b).reshape(nb, nj * 2, 1)
    b = torch.cat([b, torch.zeros(nb, 3, 1, dtype=b.dtype)], dim=1)
    w = validity_mask.float() + np.float32(1e-4)
    w = w.repeat_interleave(2, dim=1).unsqueeze(-1)
    w = torch.cat([w, torch.full((nb, 3, 1), float(np.sqrt(1e-2)), dtype=torch.float32)], dim=1)
    ref = torch.linalg.lstsq(a * w, b * w).solution
    ref = torch.cat([ref[:, :2] * scale_b, ref[:, 2:] * (scale_b / scale2d)], dim=1)
    return ref.squeeze(-1)


def reconstruct_absolute(coords2d, coords3d_rel, intrinsics, cfg: PathConfig, mix_3d_inside_fov='cfg'):
    """ptu3d.py:9-33 with the full-perspective solve (weak_perspective crashes in the reference)."""
    if cfg.weak_perspective:
        raise NotImplementedError('weak-perspective reconstruction is broken in the reference (ptu.py:30,42)')
    if mix_3d_inside_fov == 'cfg':
        mix_3d_inside_fov = cfg.mix_3d_inside_fov
    inv_k = torch.linalg.inv(intrinsics.to(coords2d.dtype))
    n2d = (to_homogeneous(coords2d) @ inv_k.transpose(1, 2))[..., :2]
    in_fov = is_within_fov(coords2d, cfg)
    ref = reconstruct_ref_fullpersp(n2d, coords3d_rel, in_fov)
    abs3 = coords3d_rel + ref[:, None]
    abs2 = to_homogeneous(n2d) * (coords3d_rel[..., 2] + ref[:, 2:3]).unsqueeze(-1)  # back_project :108-110
    if mix_3d_inside_fov is not None:
        abs2 = mix_3d_inside_fov * abs3 + (1 - mix_3d_inside_fov) * abs2
    return torch.where(in_fov[..., None], abs2, abs3)


# ----------------------------------------------------------------------------------------------
# the crop model (models/metrabs.py:47-64)
# ----------------------------------------------------------------------------------------------
def metrabs_forward(sd, spec, cfg: PathConfig, n_joints, image, intrinsics, stages=None):
    """Metrabs.forward((image, intrinsics)) -> coords3d_abs [B,J,3] fp32.  ``stages`` (dict) receives the
    stage-boundary tensors used as goldens: features, coords2d, coords3d_rel."""
    if isinstance(spec, EffNetSpec):
        features = effnet_features(sd, spec, image)
    else:
        features = spec.features(sd, image)  # TF-only backbones, oracle/port_tf_backbones.py
    coords2d, coords3d_rel = heads(sd, features, cfg, n_joints)
    out = reconstruct_absolute(coords2d, coords3d_rel, intrinsics, cfg)
    if stages is not None:
        stages.update(features=features, coords2d=coords2d, coords3d_rel=coords3d_rel)
    return out


# ----------------------------------------------------------------------------------------------
# conditioned random init + synthetic inputs.  Deterministic from the seed.
# ----------------------------------------------------------------------------------------------
def synthetic_inputs(batch, proc_side, seed=0):
    """crops U[0,1) fp32 [B,3,S,S]; intrinsics [[f,0,S/2],[0,f,S/2],[0,0,1]], f ~ U[1000,1500]."""
    g = torch.Generator().manual_seed(1000 + seed)
    crops = torch.rand(batch, 3, proc_side, proc_side, generator=g)
    f = 1000 + 500 * torch.rand(batch, generator=g)
    k = torch.zeros(batch, 3, 3)
    k[:, 0, 0] = f
    k[:, 1, 1] = f
    k[:, 0, 2] = proc_side / 2
    k[:, 1, 2] = proc_side / 2
    k[:, 2, 2] = 1
    return crops, k


def _randn(g, *shape):
    return torch.randn(*shape, generator=g)


def _calibrate_bn(sd, key, x, g, eps, gamma_scale=1.0):
    """Random affine, running stats := statistics of the calibration batch ``x`` (pre-BN activations)."""
    c = x.shape[1]
    sd[key + '.weight'] = gamma_scale * (0.8 + 0.4 * torch.rand(c, generator=g))
    sd[key + '.bias'] = 0.1 * _randn(g, c)
    sd[key + '.running_mean'] = x.mean(dim=(0, 2, 3))
    sd[key + '.running_var'] = x.var(dim=(0, 2, 3), unbiased=False) + 1e-4
    sd[key + '.num_batches_tracked'] = torch.tensor(1)
    return _bn(sd, key, x, eps)


def _init_conv_bn(sd, key, x, g, cout, k, stride=1, groups=1, act=True, eps=BN_EPS_EFFNETV2, gamma_scale=1.0):
    cin = x.shape[1] // groups
    sd[key + '.0.weight'] = _randn(g, cout, cin, k, k) * math.sqrt(2.0 / (cin * k * k))
    x = F.conv2d(x, sd[key + '.0.weight'], None, stride=stride, groups=groups)
    x = _calibrate_bn(sd, key + '.1', x, g, eps, gamma_scale)
    return F.silu(x) if act else x


def make_effnet_state_dict(spec: EffNetSpec, cfg: PathConfig, n_joints, seed=0, calib_batch=4, head_gain=10.0):
    """Conditioned random init: default init gives dead features and a degenerate LS solve, so BN
    running stats are calibrated layer by layer on a fixed random batch, residual-branch BNs are damped, and the
    head is scaled so heatmaps are peaky.  Returns a state_dict in the reference key schema."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    prefix = 'backbone.1'
    calib, _ = synthetic_inputs(calib_batch, cfg.proc_side, seed=seed + 77)
    with torch.no_grad():
        x = calib * 2 - 1
        x = _init_conv_bn(sd, f'{prefix}.0', _fixed_pad(x, 3, 0), g, spec.stem_channels, 3, stride=2)
        for b in effnet_block_list(spec):
            key = f'{prefix}.{b["key"]}.block'
            inp = x
            cexp = b['cin'] * b['expand']
            damp = 0.5 if b['residual'] else 1.0
            if b['block'] == 'fused':
                x = _fixed_pad(x, b['kernel'], b['shift'])
                if b['expand'] != 1:
                    x = _init_conv_bn(sd, f'{key}.0', x, g, cexp, b['kernel'], stride=b['stride'])
                    x = _init_conv_bn(sd, f'{key}.1', x, g, b['cout'], 1, act=False, gamma_scale=damp)
                else:
                    x = _init_conv_bn(sd, f'{key}.0', x, g, b['cout'], b['kernel'], stride=b['stride'],
                                      gamma_scale=damp)
            else:
                i = 0
                if b['expand'] != 1:
                    x = _init_conv_bn(sd, f'{key}.{i}', x, g, cexp, 1)
                    i += 1
                x = _fixed_pad(x, b['kernel'], b['shift'])
                x = _init_conv_bn(sd, f'{key}.{i}', x, g, cexp, b['kernel'], stride=b['stride'], groups=cexp)
                i += 1
                csq = max(1, b['cin'] // 4)
                sd[f'{key}.{i}.fc1.weight'] = _randn(g, csq, cexp, 1, 1) * math.sqrt(2.0 / cexp)
                sd[f'{key}.{i}.fc1.bias'] = 0.2 * _randn(g, csq)
                sd[f'{key}.{i}.fc2.weight'] = _randn(g, cexp, csq, 1, 1) * math.sqrt(2.0 / csq)
                sd[f'{key}.{i}.fc2.bias'] = 0.5 * _randn(g, cexp)
                s = x.mean(dim=(2, 3), keepdim=True)
                s = F.silu(F.conv2d(s, sd[f'{key}.{i}.fc1.weight'], sd[f'{key}.{i}.fc1.bias']))
                s = torch.sigmoid(F.conv2d(s, sd[f'{key}.{i}.fc2.weight'], sd[f'{key}.{i}.fc2.bias']))
                x = x * s
                i += 1
                x = _init_conv_bn(sd, f'{key}.{i}', x, g, b['cout'], 1, act=False, gamma_scale=damp)
            if b['residual']:
                x = x + inp
        x = _init_conv_bn(sd, f'{prefix}.{len(spec.stages) + 1}', x, g, spec.last_channel, 1)
    init_head(sd, g, spec.last_channel, n_joints, cfg.depth, head_gain)
    return sd


def init_head(sd, g, channels, n_joints, depth, head_gain=10.0):
    n_out = n_joints * (1 + depth)
    sd['heatmap_heads.conv_final.weight'] = _randn(g, n_out, channels, 1, 1) * (head_gain / math.sqrt(channels))
    sd['heatmap_heads.conv_final.bias'] = (torch.rand(n_out, generator=g) - 0.5) * 0.1


def head_only_inputs(batch, channels, hw, n_joints, depth, seed=0, head_gain=8.0):
    """Config c5 (head-only isolation): features N(0,1) rounded to bf16 so both sides see identical values,
    head weight N(0,(gain/sqrt(C))^2) rounded to bf16, bias U(-.05,.05)  (SURVEY.md 8d)."""
    g = torch.Generator().manual_seed(2000 + seed)
    feats = torch.randn(batch, channels, hw, hw, generator=g).bfloat16().float()
    sd = {}
    init_head(sd, g, channels, n_joints, depth, head_gain)
    sd['heatmap_heads.conv_final.weight'] = sd['heatmap_heads.conv_final.weight'].bfloat16().float()
    return feats, sd


def relative_error(a, b):
    """The parity metric of SURVEY.md 8d: ||a-b||_inf / ||b||_inf."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
