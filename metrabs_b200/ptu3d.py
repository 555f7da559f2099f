"""``ptu3d.reconstruct_absolute`` on the device (/root/reference/metrabs_pytorch/ptu3d.py:9-33)."""
import dataclasses

from metrabs_b200 import _lib
from metrabs_b200.engine import Engine, make_config
from metrabs_b200.util import get_config

_engines = {}


def _geometry_engine(cfg, n_joints, device_index):
    key = (cfg.proc_side, cfg.stride_train, cfg.stride_test, cfg.centered_stride, cfg.mix_3d_inside_fov, n_joints,
           device_index)
    if key not in _engines:
        c = dataclasses.replace(cfg, precision='fp32')
        n = n_joints
        # the geometry-only handle needs no weights: head-only arch with dummy channel count, depth padded so that
        # J*(1+D) is a multiple of 4
        _engines[key] = Engine(make_config(c, n, arch=_lib.ARCH_HEAD_ONLY, feature_channels=4, device=device_index))
    return _engines[key]


def reconstruct_absolute(coords2d, coords3d_rel, intrinsics, mix_3d_inside_fov=None, weak_perspective=None):
    """Same signature as the reference.  NOTE: like the reference, ``mix_3d_inside_fov=None`` means "no mixing"."""
    cfg = get_config()
    if weak_perspective is None:
        weak_perspective = cfg.weak_perspective
    if weak_perspective:
        raise NotImplementedError('weak-perspective reconstruction crashes in the reference (ptu.py:30,42)')
    if not coords2d.is_cuda:
        raise _lib.MetrabsB200Error('reconstruct_absolute needs CUDA tensors (no CPU fallback)')
    c = dataclasses.replace(cfg, mix_3d_inside_fov=mix_3d_inside_fov, depth=3)
    eng = _geometry_engine(c, coords2d.shape[1], coords2d.device.index or 0)
    return eng.reconstruct_absolute(coords2d, coords3d_rel, intrinsics)
