"""CPU: pins the oracle port (oracle/port.py) against tensors produced by the UNMODIFIED reference
(tests/golden/*.npz, written by oracle/gen_golden.py in the build container)."""
import os

import numpy as np
import pytest
import torch

from oracle import port


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_soft_argmax_matches_reference(golden_dir):
    g = _load(golden_dir, 'decode_functions.npz')
    for i in range(int(g['n_sa3d'])):
        x = torch.from_numpy(g[f'sa3d_{i}_in'])
        out = port.soft_argmax(x, dims=(4, 3, 1))
        np.testing.assert_allclose(out.numpy(), g[f'sa3d_{i}_out'], rtol=0, atol=2e-6)
    for i in range(int(g['n_sa2d'])):
        x = torch.from_numpy(g[f'sa2d_{i}_in'])
        out = port.soft_argmax(x, dims=(3, 2))
        np.testing.assert_allclose(out.numpy(), g[f'sa2d_{i}_out'], rtol=0, atol=2e-6)


def test_geometry_matches_reference(golden_dir):
    g = _load(golden_dir, 'decode_functions.npz')
    for ci, (s, st, cs, lb) in enumerate(g['geo_cfgs']):
        cfg = port.PathConfig(proc_side=int(s), stride_test=int(st), centered_stride=bool(cs),
                              legacy_centered_stride_bug=bool(lb))
        for nb, nj in [(3, 24), (1, 8), (5, 122)]:
            tag = f'geo_{ci}_{nb}_{nj}'
            u = torch.from_numpy(g[tag + '_u'])
            np.testing.assert_allclose(port.heatmap_to_image(u[..., :2], cfg).numpy(), g[tag + '_img'], rtol=1e-6)
            np.testing.assert_allclose(port.heatmap_to_metric(u, cfg).numpy(), g[tag + '_met'], rtol=1e-6)
            c2d, c3d, k = (torch.from_numpy(g[tag + n]) for n in ('_c2d', '_c3d', '_k'))
            assert (port.is_within_fov(c2d, cfg).numpy() == g[tag + '_infov']).all()
            out = port.reconstruct_absolute(c2d, c3d, k, cfg)
            assert port.relative_error(out, g[tag + '_out']) < 1e-5
            out = port.reconstruct_absolute(c2d, c3d, k, cfg, mix_3d_inside_fov=None)
            assert port.relative_error(out, g[tag + '_out_nomix']) < 1e-5


def test_weak_perspective_rejected():
    cfg = port.PathConfig(weak_perspective=True)
    with pytest.raises(NotImplementedError):
        port.reconstruct_absolute(torch.zeros(1, 2, 2), torch.zeros(1, 2, 3), torch.eye(3)[None], cfg)


def test_head_only_matches_reference(golden_dir):
    g = _load(golden_dir, 'head_only_c256_hw32_j24_d8.npz')
    cfg = port.PathConfig(proc_side=256, stride_test=8, depth=8)
    feats, sd = port.head_only_inputs(3, 256, 32, 24, 8, seed=0)
    c2d, c3d = port.heads(sd, feats, cfg, 24)
    assert port.relative_error(c2d, g['coords2d']) < 1e-5
    assert port.relative_error(c3d, g['coords3d_rel']) < 1e-5


@pytest.mark.parametrize('fname', ['tiny_s64_j8.npz', 'tiny_s128_j8_legacy.npz'])
def test_tiny_model_with_committed_weights(golden_dir, fname):
    g = _load(golden_dir, fname)
    cfg = port.PathConfig(proc_side=int(g['proc_side']), centered_stride=bool(g['centered_stride']),
                          legacy_centered_stride_bug=bool(g['legacy_centered_stride_bug']))
    spec = port.effnet_spec(str(g['name']), centered_stride=cfg.centered_stride)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd/')}
    stages = {}
    with torch.inference_mode():
        out = port.metrabs_forward(sd, spec, cfg, int(g['n_joints']), torch.from_numpy(g['crops']),
                                   torch.from_numpy(g['intrinsics']), stages=stages)
    feats = stages['features'].numpy().reshape(int(g['batch']), -1)
    assert port.relative_error(feats, g['features']) < 1e-5
    assert port.relative_error(stages['coords2d'], g['coords2d']) < 1e-5
    assert port.relative_error(stages['coords3d_rel'], g['coords3d_rel']) < 1e-5
    assert port.relative_error(out, g['coords3d_abs']) < 1e-4
    # the committed weights are what the seeded init recipe regenerates
    sd2 = port.make_effnet_state_dict(spec, cfg, int(g['n_joints']), seed=0)
    assert sd2.keys() == sd.keys()
    for k in sd:
        assert port.relative_error(sd2[k].float(), sd[k].float()) < 1e-4, k


@pytest.mark.parametrize('fname', ['effnetv2s_s256_j24.npz', 'effnetv2s_s256_j122.npz',
                                   'effnetv2l_s256_j24.npz', 'effnetv2l_s384_j24.npz'])
def test_full_models_regenerated_from_seed(golden_dir, fname):
    """Weights are regenerated from the seed (too large to commit); the init includes a BN calibration forward,
    so cross-machine float summation order leaves ~1e-5 noise - tolerance 1e-3 like the device parity bar."""
    g = _load(golden_dir, fname)
    s, j, b = int(g['proc_side']), int(g['n_joints']), int(g['batch'])
    cfg = port.PathConfig(proc_side=s)
    spec = port.effnet_spec(str(g['name']))
    sd = port.make_effnet_state_dict(spec, cfg, j, seed=0)
    crops, k = port.synthetic_inputs(b, s, seed=0)
    stages = {}
    with torch.inference_mode():
        out = port.metrabs_forward(sd, spec, cfg, j, crops, k, stages=stages)
    feats = stages['features'].numpy().reshape(b, -1)[:, ::int(g['feature_stride'])]
    assert port.relative_error(feats, g['features']) < 1e-3
    assert port.relative_error(stages['coords2d'], g['coords2d']) < 1e-3
    assert port.relative_error(out, g['coords3d_abs']) < 1e-3
