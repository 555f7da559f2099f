"""heatmap [0,1] coordinates -> image pixels / metric mm (/root/reference/metrabs_pytorch/models/util.py:6-33).
On the product path these affine maps are applied inside the decode kernel epilogue (DecodeScale in
csrc/decode.cuh); the functions here exist for API parity and operate on whatever device the input lives on."""
import torch

from metrabs_b200.util import get_config


def heatmap_to_image(coords, is_training=False):
    cfg = get_config()
    stride = cfg.stride_train if is_training else cfg.stride_test
    last_image_pixel = cfg.proc_side - 1
    last_receptive_center = last_image_pixel - (last_image_pixel % stride)
    out = coords * last_receptive_center
    if cfg.centered_stride:
        out = out + stride // 2
    if cfg.legacy_centered_stride_bug:
        out = out + stride // 2
    return out


def heatmap_to_25d(coords, is_training=False):
    cfg = get_config()
    return torch.cat([heatmap_to_image(coords[..., :2], is_training), coords[..., 2:] * cfg.box_size_mm], dim=-1)


def heatmap_to_metric(coords, is_training=False):
    cfg = get_config()
    xy = heatmap_to_image(coords[..., :2], is_training) * cfg.box_size_mm / cfg.proc_side
    return torch.cat([xy, coords[..., 2:] * cfg.box_size_mm], dim=-1)
