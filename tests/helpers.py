"""Shared builders for the GPU parity tests: the oracle side (oracle/port.py state_dict + forward) and the device
side (metrabs_b200.Metrabs behind the C ABI) on identical weights and inputs."""
import dataclasses
import types

import torch

import metrabs_b200
from metrabs_b200.backbones import efficientnet as E
from metrabs_b200.models.metrabs import Metrabs
from oracle import port

SIZES = {'efficientnetv2-tiny': 'tiny', 'efficientnetv2-s': 's', 'efficientnetv2-m': 'm', 'efficientnetv2-l': 'l'}


def joint_info(n):
    return types.SimpleNamespace(names=[f'j{i}' for i in range(n)], stick_figure_edges=[(0, 1)], n_joints=n)


def device_model(name, pcfg: port.PathConfig, n_joints, sd, precision='fp32'):
    cfg = metrabs_b200.Config(**{k: v for k, v in dataclasses.asdict(pcfg).items()}, precision=precision)
    metrabs_b200.set_config(cfg)
    bb = E.EfficientNet(SIZES[name])
    m = Metrabs(torch.nn.Sequential(E.PreprocLayer(), bb.features), joint_info(n_joints)).eval()
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def rel_err(a, b):
    return port.relative_error(a.detach().float().cpu(), torch.as_tensor(b).float().cpu())


def device_model_tf(kind, pcfg: port.PathConfig, n_joints, sd, precision='fp32'):
    """kind: 'resnet50' | 'mobilenetv3-small' (TF-only backbones; key schema defined by this build)."""
    from metrabs_b200.backbones import mobilenet_v3, resnet
    cfg = metrabs_b200.Config(**{k: v for k, v in dataclasses.asdict(pcfg).items()}, precision=precision)
    metrabs_b200.set_config(cfg)
    feats = resnet.resnet50() if kind == 'resnet50' else mobilenet_v3.mobilenet_v3_small()
    m = Metrabs(feats, joint_info(n_joints)).eval()
    m.load_state_dict(sd, strict=True)
    return m.cuda()
