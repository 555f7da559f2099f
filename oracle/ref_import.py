"""TEST INFRASTRUCTURE ONLY - loader for the *unmodified* reference (isarandi/metrabs) from /root/reference.

Only usable in the build container (``/root/reference`` does not exist on the GPU box).  It is used by
``oracle/gen_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by the CPU tests that
pin the oracle port (``oracle/port.py``) against the real reference when the reference tree is present.

The reference imports ``hydra``, ``posepile`` and ``simplepyutils`` (not installed, no network); empty stub
modules are injected and ``metrabs_pytorch.util._cfg`` is set directly, which bypasses hydra
(``/root/reference/metrabs_pytorch/util.py:41-57``).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('METRABS_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'metrabs_pytorch'))


def _stub(name, **attrs):
    if name in sys.modules:
        mod = sys.modules[name]
    else:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def import_reference(cfg_dict):
    """Returns the reference package modules with its global config set to ``cfg_dict``."""
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _stub('hydra')
    _stub('hydra.core')
    _stub('hydra.core.global_hydra')
    _stub('posepile')
    _stub('posepile.paths', DATA_ROOT='/tmp')
    _stub('posepile.datasets3d')
    _stub('simplepyutils', FLAGS=types.SimpleNamespace())
    import metrabs_pytorch.util as ref_util
    ref_util._cfg = types.SimpleNamespace(**cfg_dict)
    ref_util.get_config.cache_clear()
    import metrabs_pytorch.ptu as ref_ptu
    import metrabs_pytorch.ptu3d as ref_ptu3d
    import metrabs_pytorch.models.util as ref_model_util
    import metrabs_pytorch.models.metrabs as ref_metrabs
    import metrabs_pytorch.backbones.efficientnet as ref_effnet
    return types.SimpleNamespace(
        util=ref_util, ptu=ref_ptu, ptu3d=ref_ptu3d, model_util=ref_model_util,
        metrabs=ref_metrabs, effnet=ref_effnet)


def set_reference_config(cfg_dict):
    import metrabs_pytorch.util as ref_util
    ref_util._cfg = types.SimpleNamespace(**cfg_dict)
    ref_util.get_config.cache_clear()
