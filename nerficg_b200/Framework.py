"""Configuration and runtime shell with the reference's interface (reference src/Framework.py).

Only what the NeRF method plugin needs: the global ``config`` (attribute-access dict loaded from
the same YAML files / ``KEY=VAL`` overrides), the ``Configurable.configure`` class decorator that
copies UPPERCASE defaults + config-section overrides onto instances (Framework.py:73-108), the
exception hierarchy (Framework.py:360-428), seeding and device setup.  There is no CPU mode here:
``setup_torch`` selects ``cuda:GPU_INDICES[0]`` or raises.

Drop-in mode: the reference keeps its configuration in the module global ``Framework.config`` and REBINDS it in
``load_config`` (Framework.py:163-176), and every ``Configurable`` reads that global at construction
(Framework.py:73-88).  ``bind_host_framework(reference_Framework_module)`` (called by
``Implementations.install_into_reference``) makes ``nerficg_b200.Framework.config`` / ``.Directories`` resolve to
the host module's objects at every access (module ``__getattr__``), so the host run's YAML sections,
``KEY=VAL`` overrides, ``GLOBAL.DEFAULT_DEVICE`` and output directories are the ones our plugin classes see.
"""
from __future__ import annotations

import ast
import random
from pathlib import Path
from typing import Any

import numpy as np
import torch
import yaml

from .Logging import Logger


class ConfigParameterList(dict):
    """dict with attribute access and recursive update (stands in for Munch; Framework.py:39-53)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]

    def copy(self) -> 'ConfigParameterList':
        return type(self)({k: (v.copy() if isinstance(v, ConfigParameterList) else v) for k, v in self.items()})

    @classmethod
    def fromDict(cls, d):
        if isinstance(d, dict):
            return cls({k: cls.fromDict(v) for k, v in d.items()})
        if isinstance(d, list):
            return [cls.fromDict(v) for v in d]
        return d

    def toDict(self) -> dict:
        return {k: (v.toDict() if isinstance(v, ConfigParameterList) else v) for k, v in self.items()}

    def recursive_update(self, other: 'ConfigParameterList') -> None:
        if not isinstance(other, ConfigParameterList):
            raise TypeError()
        for key, value in other.items():
            if isinstance(value, ConfigParameterList) and isinstance(self.get(key), ConfigParameterList):
                self[key].recursive_update(value)
            else:
                self[key] = value.copy() if isinstance(value, ConfigParameterList) else value


class _LocalDirectories:
    ROOT: Path = Path(__file__).resolve().parent.parent
    NERFICG_ROOT: Path = ROOT
    OUTPUT_DIR: Path = ROOT / 'output'
    CONFIG_DIR: Path = ROOT / 'configs'


def get_default_global_config() -> ConfigParameterList:
    return ConfigParameterList(LOG_LEVEL=Logger.MODE_VERBOSE, GPU_INDICES=[0], RANDOM_SEED=0, ANOMALY_DETECTION=False,
                               FILTER_WARNINGS=True, METHOD_TYPE='NeRF', DATASET_TYPE='NeRF')


_local_config: ConfigParameterList = ConfigParameterList(GLOBAL=get_default_global_config(),
                                                         TRAINING=ConfigParameterList(WANDB=ConfigParameterList(ACTIVATE=False)))
_host = None  # the reference's Framework module once bound (drop-in mode)


def bind_host_framework(host_framework_module) -> None:
    """Drop-in mode: resolve ``config`` and ``Directories`` through the host framework's module from now on."""
    global _host
    if not hasattr(host_framework_module, 'Configurable') or not hasattr(host_framework_module, 'Directories'):
        raise FrameworkError(f'{host_framework_module!r} is not a nerficg Framework module')
    _host = host_framework_module


def unbind_host_framework() -> None:
    global _host
    _host = None


def host_framework():
    return _host


def current_config():
    """The live configuration: the host framework's ``config`` global when bound (looked up at every call because the
    host rebinds it in load_config / deletes it in teardown), else this module's own."""
    if _host is not None:
        cfg = getattr(_host, 'config', None)
        if cfg is None:
            raise FrameworkError('host framework is bound but has no config loaded (call its Framework.setup() first)')
        return cfg
    return _local_config


def __getattr__(name: str):  # PEP 562: Framework.config / Framework.Directories are resolved at access time
    if name == 'config':
        return current_config()
    if name == 'Directories':
        return _host.Directories if _host is not None else _LocalDirectories
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')


class Configurable:
    """Mixin: class-level defaults declared by ``configure`` are merged base -> derived, overridden by the
    config-file section named at construction and copied onto the instance as plain attributes."""
    _configuration: ConfigParameterList = ConfigParameterList()

    def __init__(self, config_file_data_field: str) -> None:
        self.config_file_data_field = config_file_data_field
        params = type(self)._configuration.copy()
        cfg = current_config()
        section = cfg.get(config_file_data_field) if hasattr(cfg, 'get') else getattr(cfg, config_file_data_field, None)
        if section is None:
            Logger.log_debug(f'config section {config_file_data_field} missing for {type(self).__name__}: using defaults')
        else:
            # the host framework's sections are Munch objects: converted (recursively) to our container
            params.recursive_update(section if isinstance(section, ConfigParameterList) else ConfigParameterList.fromDict(dict(section)))
        for key in params:
            self.__dict__[key] = params[key]

    @classmethod
    def get_default_parameters(cls) -> ConfigParameterList:
        return cls._configuration

    @staticmethod
    def configure(**params):
        new_params = ConfigParameterList(params)

        def decorator(cls):
            if not issubclass(cls, Configurable):
                raise FrameworkError(f'configure decorator must be applied to a subclass of Configurable, got {cls}')
            merged = ConfigParameterList()
            for base in cls.__bases__:
                if issubclass(base, Configurable):
                    merged.recursive_update(base._configuration)
            merged.recursive_update(new_params)
            cls._configuration = merged
            return cls
        return decorator


def load_config(config_path: Path | str | None, overrides: dict[str, Any] | None = None) -> ConfigParameterList:
    """YAML -> global ``config``; ``overrides`` are dotted ``A.B.C`` keys (strings are literal_eval'ed),
    the reference's ``-c cfg.yaml KEY=VAL`` mechanism (Framework.py:163-199)."""
    global _local_config
    if _host is not None:
        raise FrameworkError('a host framework is bound: load the configuration through the host\'s Framework.setup()')
    if config_path is not None:
        with open(config_path) as f:
            config = ConfigParameterList.fromDict(yaml.safe_load(f))
        config.path = Path(config_path)
    else:
        config = ConfigParameterList(GLOBAL=get_default_global_config())
    _local_config = config
    defaults = get_default_global_config()
    config.setdefault('GLOBAL', ConfigParameterList())
    for k, v in defaults.items():
        config.GLOBAL.setdefault(k, v)
    for section in ('MODEL', 'RENDERER', 'TRAINING', 'DATASET'):
        config.setdefault(section, ConfigParameterList())
    config.TRAINING.setdefault('WANDB', ConfigParameterList(ACTIVATE=False))
    for dotted, value in (overrides or {}).items():
        if isinstance(value, str):
            try:
                value = ast.literal_eval(value)
            except (ValueError, SyntaxError):
                pass
        *path, leaf = dotted.split('.')
        target = config
        for key in path:
            if key not in target:
                raise FrameworkError(f'invalid config key "{key}" in override "{dotted}"')
            target = target[key]
        target[leaf] = value
    Logger.set_mode(config.GLOBAL.LOG_LEVEL)
    return config


def set_random_seed() -> None:
    config = current_config()
    if config.GLOBAL.RANDOM_SEED is None:
        config.GLOBAL.RANDOM_SEED = int(np.random.randint(0, 2 ** 31 - 1))
    torch.manual_seed(config.GLOBAL.RANDOM_SEED)
    random.seed(config.GLOBAL.RANDOM_SEED)
    np.random.seed(config.GLOBAL.RANDOM_SEED)


def setup_torch(device_index: int | None = None) -> torch.device:
    """Selects the CUDA device (``GPU_INDICES[0]`` unless given).  No CPU mode: raises without a GPU."""
    from . import _lib
    if not torch.cuda.is_available():
        raise FrameworkError('nerficg_b200 needs a B200 GPU: CUDA is not available and there is no CPU path')
    config = current_config()
    if device_index is None:
        indices = config.GLOBAL.GPU_INDICES or [0]
        device_index = indices[0]
    device = torch.device(f'cuda:{device_index}')
    torch.cuda.set_device(device)
    _lib.check(_lib.load().nerf_device_check(device_index), 'nerf_device_check')
    config.GLOBAL.DEFAULT_DEVICE = device
    return device


def setup(config_path: Path | str | None = None, overrides: dict[str, Any] | None = None, device_index: int | None = None):
    load_config(config_path, overrides)
    device = setup_torch(device_index)
    set_random_seed()
    return device


# ---- exception hierarchy (names as in the reference) -------------------------------------
class FrameworkError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        Logger.log_error(f'({self.__class__.__name__}) {msg}')


class MethodError(FrameworkError):
    pass


class CheckpointError(FrameworkError):
    pass


class RendererError(FrameworkError):
    pass


class ModelError(FrameworkError):
    pass


class TrainingError(FrameworkError):
    pass


class CameraError(FrameworkError):
    pass


class DatasetError(FrameworkError):
    pass


class LossError(FrameworkError):
    pass


class SamplerError(FrameworkError):
    pass
