"""Imports the UNMODIFIED reference (/root/reference) in the build container.  TEST INFRASTRUCTURE ONLY.

Used by ``oracle/make_golden.py`` (and, when the tree is present, by the
``--impl reference`` arm of ``bench.py``) to run nerficg's own NeRF code on CPU.
The reference needs four packages that are not installed here (munch, natsort,
plyfile, torchmetrics); none of them does arithmetic on the hot path except
torchmetrics' PSNR, so tiny in-memory stand-ins are registered before import
(SURVEY.md Appendix C).  Nothing from the reference is copied into this repo.
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

REFERENCE_ROOTS = [Path('/root/reference'), Path(__file__).resolve().parent.parent / 'baseline' / '_ref']


def reference_root() -> Path | None:
    for root in REFERENCE_ROOTS:
        if (root / 'src' / 'Methods' / 'NeRF' / 'Renderer.py').exists():
            return root
    return None


def _install_stubs() -> None:
    import torch

    if 'munch' not in sys.modules:
        m = types.ModuleType('munch')

        class Munch(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v

            def __delattr__(self, k):
                del self[k]

            def copy(self):
                return type(self)(self)

            @classmethod
            def fromDict(cls, d):
                if isinstance(d, dict):
                    return cls({k: cls.fromDict(v) for k, v in d.items()})
                if isinstance(d, list):
                    return [cls.fromDict(v) for v in d]
                return d

            def toDict(self):
                return {k: (v.toDict() if isinstance(v, Munch) else v) for k, v in self.items()}

        m.Munch = Munch
        sys.modules['munch'] = m
    if 'natsort' not in sys.modules:
        m = types.ModuleType('natsort')
        m.natsorted = sorted
        sys.modules['natsort'] = m
    if 'plyfile' not in sys.modules:
        m = types.ModuleType('plyfile')
        m.PlyData = object
        m.PlyElement = object
        sys.modules['plyfile'] = m
    if 'torchmetrics' not in sys.modules:
        tm = types.ModuleType('torchmetrics')
        tm.Metric = type('Metric', (), {})
        fn = types.ModuleType('torchmetrics.functional')
        fi = types.ModuleType('torchmetrics.functional.image')

        def peak_signal_noise_ratio(preds, target, data_range=1.0):
            return 10.0 * torch.log10(data_range ** 2 / torch.mean((preds - target) ** 2))

        fi.peak_signal_noise_ratio = peak_signal_noise_ratio
        fn.image = fi
        tm.functional = fn
        im = types.ModuleType('torchmetrics.image')
        tm.image = im
        sys.modules.update({'torchmetrics': tm, 'torchmetrics.functional': fn,
                            'torchmetrics.functional.image': fi, 'torchmetrics.image': im})


_loaded = None


def load_reference(n_samples: int = 192, coarse_ratio: float = 0.3333333, seed: int = 0):
    """Returns (Framework, NeRF method module, utils module, Datasets.utils, Cameras) of the reference in CPU mode."""
    global _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError('reference tree not present (expected /root/reference)')
    _install_stubs()
    src = str(root / 'src')
    if src not in sys.path:
        sys.path.insert(0, src)
    import torch
    import warnings
    warnings.filterwarnings('ignore')
    cwd = os.getcwd()
    import Framework  # noqa: the reference's module
    from Logging import Logger
    Framework.load_config(root / 'configs' / 'nerf_lego.yaml', True,
                          {'RENDERER.N_SAMPLES': str(n_samples), 'RENDERER.COARSE_RATIO': str(coarse_ratio),
                           'GLOBAL.RANDOM_SEED': str(seed)})
    Framework.config.GLOBAL.GPU_INDICES = None
    Framework.config.GLOBAL.LOG_LEVEL = 0
    Framework.config.TRAINING.WANDB.ACTIVATE = False
    Logger.set_mode(0)
    Framework.setup_torch()
    Framework.set_random_seed()
    import Implementations
    method = Implementations.Methods.import_method('NeRF')
    import Methods.NeRF.utils as nerf_utils
    import Datasets.utils as ds_utils
    from Cameras.Perspective import PerspectiveCamera
    from Cameras.utils import SharedCameraSettings
    os.chdir(cwd)
    _loaded = dict(Framework=Framework, method=method, utils=nerf_utils, ds_utils=ds_utils,
                   PerspectiveCamera=PerspectiveCamera, SharedCameraSettings=SharedCameraSettings)
    torch.set_default_dtype(torch.float32)
    return _loaded
