"""TEST INFRASTRUCTURE ONLY - generates the committed golden fixtures under tests/golden/ by running the
UNMODIFIED reference (imported from /root/reference via oracle/ref_import.py) on torch-cpu.

Run in the build container only:  ``python oracle/gen_golden.py``.  The GPU box never runs this (it has no
/root/reference); it only reads the .npz files.  Weights come from ``oracle/port.make_effnet_state_dict``
(conditioned random init, deterministic from the seed) loaded into the reference model with
``load_state_dict(strict=True)``, so the reference key schema is exercised too.
"""
import os
import sys
import types
from functools import partial

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port  # noqa: E402
from oracle.ref_import import import_reference, set_reference_config  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def build_reference_model(R, spec, n_joints, proc_side):
    E = R.effnet
    rows = []
    for st in spec.stages:
        C = E.FusedMBConvConfig if st.block == 'fused' else E.MBConvConfig
        rows.append(C(st.expand, st.kernel, st.stride, st.cin, st.cout, st.layers,
                      bottomright_stride=st.bottomright))
    bb = E.EfficientNet(rows, 0.2, last_channel=spec.last_channel,
                        norm_layer=partial(torch.nn.BatchNorm2d, eps=1e-3))
    ji = types.SimpleNamespace(names=[f'j{i}' for i in range(n_joints)], stick_figure_edges=[(0, 1)],
                               n_joints=n_joints)
    m = R.metrabs.Metrabs(torch.nn.Sequential(E.PreprocLayer(), bb.features), ji).eval()
    with torch.inference_mode():  # materialise LazyConv2d (scripts/demo_image.py:69-72)
        m((torch.rand(1, 3, proc_side, proc_side), torch.eye(3)[None]))
    return m


def state_dict_checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items()) if v.ndim > 0))


def model_golden(R, name, proc_side, n_joints, batch, fname, store_weights=False, feature_stride=1,
                 centered_stride=True, legacy_bug=False):
    cfg = port.PathConfig(proc_side=proc_side, centered_stride=centered_stride,
                          legacy_centered_stride_bug=legacy_bug)
    set_reference_config(cfg.as_reference_dict())
    spec = port.effnet_spec(name, centered_stride=centered_stride)
    sd = port.make_effnet_state_dict(spec, cfg, n_joints, seed=0)
    m = build_reference_model(R, spec, n_joints, proc_side)
    m.load_state_dict(sd, strict=True)
    crops, k = port.synthetic_inputs(batch, proc_side, seed=0)
    with torch.inference_mode():
        feats = m.backbone(crops)
        c2d, c3d = m.heatmap_heads(feats)
        out = m((crops, k))
    data = dict(
        name=name, proc_side=proc_side, n_joints=n_joints, batch=batch, seed=0,
        centered_stride=centered_stride, legacy_centered_stride_bug=legacy_bug,
        feature_stride=feature_stride,
        state_dict_checksum=state_dict_checksum(sd),
        features=feats.numpy().reshape(batch, -1)[:, ::feature_stride].copy(),
        features_absmean=float(feats.abs().mean()),
        coords2d=c2d.numpy(), coords3d_rel=c3d.numpy(), coords3d_abs=out.numpy())
    if store_weights:
        data['crops'] = crops.numpy()
        data['intrinsics'] = k.numpy()
        for key, v in sd.items():
            data['sd/' + key] = v.numpy()
    np.savez_compressed(os.path.join(OUT, fname), **data)
    print(fname, 'abs range', float(out.min()), float(out.max()), 'checksum', data['state_dict_checksum'])


def decode_goldens(R):
    """Per-function goldens for ptu.soft_argmax, models/util.heatmap_to_*, ptu3d.reconstruct_absolute."""
    g = torch.Generator().manual_seed(123)
    data = {}
    # soft_argmax, 3D (dims (4,3,1) on [B,D,J,H,W]) and 2D (dims (3,2) on [B,J,H,W]); incl. size-1 axes (0.5 rule)
    shapes3d = [(2, 8, 5, 8, 8), (1, 8, 24, 12, 12), (2, 1, 3, 4, 4), (1, 4, 2, 1, 6), (1, 32, 2, 32, 32)]
    for i, shp in enumerate(shapes3d):
        x = torch.randn(*shp, generator=g) * 4
        data[f'sa3d_{i}_in'] = x.numpy()
        data[f'sa3d_{i}_out'] = R.ptu.soft_argmax(x, dim=(4, 3, 1)).numpy()
    shapes2d = [(2, 5, 8, 8), (1, 24, 12, 12), (2, 3, 1, 7), (1, 2, 32, 32)]
    for i, shp in enumerate(shapes2d):
        x = torch.randn(*shp, generator=g) * 4
        data[f'sa2d_{i}_in'] = x.numpy()
        data[f'sa2d_{i}_out'] = R.ptu.soft_argmax(x, dim=(3, 2)).numpy()
    data['n_sa3d'] = len(shapes3d)
    data['n_sa2d'] = len(shapes2d)
    # heatmap scaling + reconstruct_absolute under several configs
    cfgs = [dict(proc_side=256, stride_test=32, centered_stride=True, legacy_centered_stride_bug=False),
            dict(proc_side=384, stride_test=32, centered_stride=True, legacy_centered_stride_bug=False),
            dict(proc_side=256, stride_test=32, centered_stride=False, legacy_centered_stride_bug=True),
            dict(proc_side=256, stride_test=8, centered_stride=True, legacy_centered_stride_bug=False),
            dict(proc_side=256, stride_test=4, centered_stride=False, legacy_centered_stride_bug=False)]
    for ci, c in enumerate(cfgs):
        cfg = port.PathConfig(**c)
        set_reference_config(cfg.as_reference_dict())
        for nb, nj in [(3, 24), (1, 8), (5, 122)]:
            u = torch.rand(nb, nj, 3, generator=g)
            img = R.model_util.heatmap_to_image(u[..., :2], False)
            met = R.model_util.heatmap_to_metric(u, False)
            s = cfg.proc_side
            # 2D points partly outside the FOV band so both branches of the final where() are hit
            c2d = torch.rand(nb, nj, 2, generator=g) * (s * 1.2) - 0.1 * s
            c3d = torch.randn(nb, nj, 3, generator=g) * torch.tensor([300., 400., 250.])
            f = 1000 + 500 * torch.rand(nb, generator=g)
            k = torch.zeros(nb, 3, 3)
            k[:, 0, 0] = f
            k[:, 1, 1] = f * 1.02
            k[:, 0, 1] = 0.5
            k[:, 0, 2] = s / 2 + 3
            k[:, 1, 2] = s / 2 - 2
            k[:, 2, 2] = 1
            # make 2D and 3D roughly consistent so that the LS problem is meaningful
            root = torch.tensor([50., -80., 3500.])
            p = c3d + root
            proj = p[..., :2] / p[..., 2:]
            c2d_cons = torch.einsum('bjk,bik->bji', torch.cat([proj, torch.ones_like(proj[..., :1])], -1), k)[..., :2]
            c2d = torch.where(torch.rand(nb, nj, 1, generator=g) < 0.8, c2d_cons, c2d)
            out = R.ptu3d.reconstruct_absolute(c2d, c3d, k, mix_3d_inside_fov=cfg.mix_3d_inside_fov)
            out_nomix = R.ptu3d.reconstruct_absolute(c2d, c3d, k, mix_3d_inside_fov=None)
            tag = f'geo_{ci}_{nb}_{nj}'
            data[tag + '_u'] = u.numpy()
            data[tag + '_img'] = img.numpy()
            data[tag + '_met'] = met.numpy()
            data[tag + '_c2d'] = c2d.numpy()
            data[tag + '_c3d'] = c3d.numpy()
            data[tag + '_k'] = k.numpy()
            data[tag + '_out'] = out.numpy()
            data[tag + '_out_nomix'] = out_nomix.numpy()
            data[tag + '_infov'] = R.ptu3d.is_within_fov(c2d).numpy()
    data['geo_cfgs'] = np.array([[c['proc_side'], c['stride_test'], int(c['centered_stride']),
                                  int(c['legacy_centered_stride_bug'])] for c in cfgs])
    np.savez_compressed(os.path.join(OUT, 'decode_functions.npz'), **data)
    print('decode_functions.npz', len(data), 'arrays')


def head_golden(R):
    """Head-only (config c5-like, small): features -> MetrabsHeads of the reference."""
    cfg = port.PathConfig(proc_side=256, stride_test=8, depth=8)
    set_reference_config(cfg.as_reference_dict())
    feats, sd = port.head_only_inputs(3, 256, 32, 24, 8, seed=0)
    hm = R.metrabs.MetrabsHeads(n_points=24).eval()
    with torch.inference_mode():
        hm(feats[:1])
        hm.conv_final.weight.copy_(sd['heatmap_heads.conv_final.weight'])
        hm.conv_final.bias.copy_(sd['heatmap_heads.conv_final.bias'])
        c2d, c3d = hm(feats)
    np.savez_compressed(os.path.join(OUT, 'head_only_c256_hw32_j24_d8.npz'), coords2d=c2d.numpy(),
                        coords3d_rel=c3d.numpy())
    print('head_only ok')


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    R = import_reference(port.PathConfig().as_reference_dict())
    decode_goldens(R)
    head_golden(R)
    model_golden(R, 'efficientnetv2-tiny', 64, 8, 3, 'tiny_s64_j8.npz', store_weights=True)
    model_golden(R, 'efficientnetv2-tiny', 128, 8, 2, 'tiny_s128_j8_legacy.npz', store_weights=True,
                 centered_stride=False, legacy_bug=True)
    model_golden(R, 'efficientnetv2-s', 256, 24, 2, 'effnetv2s_s256_j24.npz', feature_stride=16)
    model_golden(R, 'efficientnetv2-s', 256, 122, 2, 'effnetv2s_s256_j122.npz', feature_stride=16)
    model_golden(R, 'efficientnetv2-l', 256, 24, 2, 'effnetv2l_s256_j24.npz', feature_stride=16)
    model_golden(R, 'efficientnetv2-l', 384, 24, 1, 'effnetv2l_s384_j24.npz', feature_stride=16)


if __name__ == '__main__':
    main()
