"""The crop model: host-side mirror of /root/reference/metrabs_pytorch/models/metrabs.py (``Metrabs`` :12-64,
``MetrabsHeads`` :67-85) whose forward runs in libmetrabs_b200.so.

Drop-in contract: ``Metrabs(backbone, joint_info)``; ``forward((image[B,3,S,S] fp32 in [0,1],
intrinsics[B,3,3])) -> coords3d_abs[B,J,3] fp32``; attributes ``joint_names``, ``joint_edges``,
``input_resolution``, ``joint_info``, ``heatmap_heads``; ``load_state_dict`` with the reference key schema
(``backbone.1.<stage>...``, ``heatmap_heads.conv_final.{weight,bias}``).  The consumer is
``Pose3dEstimator._predict_single_batch`` (multiperson/multiperson_model.py:240-242).
"""
import numpy as np
import torch
from torch import nn

from metrabs_b200 import _lib
from metrabs_b200.engine import Engine, make_config
from metrabs_b200.util import get_config


def _find_features(backbone):
    for m in backbone.modules():
        if hasattr(m, 'arch') and hasattr(m, 'last_channel') and hasattr(m, 'stages'):
            return m
    raise TypeError('backbone must contain a metrabs_b200.backbones.*.Features module '
                    '(e.g. Sequential(PreprocLayer(), efficientnet_v2_s().features))')


class MetrabsHeads(nn.Module):
    """1x1 conv (J + D*J channels, bias) + 2D / volumetric soft-argmax + metric scaling, fused on the device."""

    def __init__(self, n_points, in_channels, owner=None):
        super().__init__()
        cfg = get_config()
        self.n_points = n_points
        self.n_outs = [n_points, cfg.depth * n_points]
        self.conv_final = nn.Conv2d(in_channels, sum(self.n_outs), kernel_size=1)
        self._owner = [owner]  # list: keep the parent out of the module tree

    def forward(self, inp):
        """features NCHW [B,C,H,W] (reference layout) -> (coords2d [B,J,2] px, coords3d_rel [B,J,3] mm)."""
        eng = self._owner[0].engine()
        nhwc = inp.permute(0, 2, 3, 1).contiguous().to(eng.feature_dtype)
        return eng.head_decode(nhwc)


class Metrabs(nn.Module):
    def __init__(self, backbone, joint_info):
        super().__init__()
        cfg = get_config()
        if cfg.affine_weights or cfg.transform_coords or cfg.predict_all_and_latents:
            # the reference's PT forward calls an undefined latent_points_to_joints for these (metrabs.py:62)
            raise NotImplementedError('affine_weights / transform_coords / predict_all_and_latents are not '
                                      'functional in the reference PyTorch path')
        self.backbone = backbone
        self.joint_names = np.array(joint_info.names)
        self.joint_edges = np.array([[i, j] for i, j in joint_info.stick_figure_edges])
        self.input_resolution = np.int32(cfg.proc_side)
        self.joint_info = joint_info
        self._features = [_find_features(backbone)]
        feats = self._features[0]
        self.heatmap_heads = MetrabsHeads(n_points=joint_info.n_joints, in_channels=feats.last_channel, owner=self)
        self._cfg = cfg
        self._engine = None
        self._dirty = True
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.mark_weights_changed())

    def mark_weights_changed(self):
        self._dirty = True

    def engine(self, device=None):
        """Builds the C handle lazily on the module's CUDA device and (re)uploads the weights when they changed."""
        if device is None:
            device = self.heatmap_heads.conv_final.weight.device
        if device.type != 'cuda':
            raise _lib.MetrabsB200Error('metrabs_b200.Metrabs runs on CUDA only: call .cuda() first (no CPU fallback)')
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._engine is None or self._engine.cfg.device != index:
            feats = self._features[0]
            self._engine = Engine(make_config(self._cfg, self.joint_info.n_joints, stages=feats.stages,
                                              last_channel=feats.last_channel, arch=feats.arch, device=index))
            self._dirty = True
        if self._dirty:
            self._engine.load_state_dict(self.state_dict())
            self._dirty = False
        return self._engine

    def forward(self, inp):
        image, intrinsics = inp
        eng = self.engine(image.device)
        return eng.forward(image.float(), intrinsics.float())
