"""Plugin registry with the reference's accessors (src/Implementations.py:18-65)."""
from __future__ import annotations

import importlib
from types import ModuleType

from . import Framework
from .Logging import Logger


class Methods:
    options = ('NeRF',)
    modules: dict[str, ModuleType] = {}

    @staticmethod
    def import_method(method: str) -> ModuleType:
        if method not in Methods.options:
            raise Framework.MethodError(f'requested invalid method type: {method}\\navailable methods are: {Methods.options}')
        if method not in Methods.modules:
            Methods.modules[method] = importlib.import_module(f'{__package__}.Methods.{method}')
        return Methods.modules[method]

    @staticmethod
    def get_model(method: str, checkpoint: str = None, name: str = 'Default'):
        Logger.log_info('creating model')
        cls = Methods.import_method(method).MODEL
        model = cls.load(checkpoint) if checkpoint is not None else cls(name).build()
        device = Framework.config.GLOBAL.get('DEFAULT_DEVICE')
        return model.to(device) if device is not None else model

    @staticmethod
    def get_renderer(method: str, model):
        Logger.log_info('creating renderer')
        return Methods.import_method(method).RENDERER(model)

    @staticmethod
    def get_training_instance(method: str, checkpoint: str | None = None, **kwargs):
        Logger.log_info('creating training instance')
        if checkpoint is not None:  # '.train' resume (reference src/Implementations.py:61-62, Base/Trainer.py:94-104)
            return Methods.import_method(method).TRAINING_INSTANCE.load(checkpoint)
        model = Methods.get_model(method, name=Framework.config.TRAINING.get('MODEL_NAME', 'Default'))
        renderer = Methods.get_renderer(method, model)
        return Methods.import_method(method).TRAINING_INSTANCE(model=model, renderer=renderer, **kwargs)


class Datasets:
    """Dataset registry (reference src/Implementations.py:68-98): ``GLOBAL.DATASET_TYPE`` -> ``CustomDataset`` class."""
    options = ('NeRF',)
    loaded: dict[str, type] = {}

    @staticmethod
    def get_dataset_class(dataset_type: str) -> type:
        if dataset_type not in Datasets.options:
            raise Framework.DatasetError(f'requested invalid dataset type: {dataset_type}\navailable datasets are: {Datasets.options}')
        if dataset_type not in Datasets.loaded:
            Datasets.loaded[dataset_type] = importlib.import_module(f'{__package__}.Datasets.{dataset_type}').CustomDataset
        return Datasets.loaded[dataset_type]

    @staticmethod
    def get_dataset(dataset_type: str, path: str):
        Logger.log_info('loading dataset')
        return Datasets.get_dataset_class(dataset_type)(path)


def install_into_reference(reference_implementations_module, reference_framework_module=None) -> None:
    """Drop-in injection: makes the UNMODIFIED nerficg scripts use these classes for METHOD_TYPE 'NeRF'.

    1. binds ``nerficg_b200.Framework.config`` / ``Directories`` to the host's ``Framework`` module, so our
       ``Configurable`` classes read the host run's YAML sections and ``KEY=VAL`` overrides, ``GLOBAL.DEFAULT_DEVICE``
       and output directories (reference src/Framework.py:73-108,163-199,263-290);
    2. pre-seeds the class-level module cache (src/Implementations.py:22,31-40) so that ``get_model`` /
       ``get_renderer`` / ``get_training_instance`` resolve METHOD_TYPE 'NeRF' to this package.
    """
    if reference_framework_module is None:
        reference_framework_module = getattr(reference_implementations_module, 'Framework', None)
    if reference_framework_module is None:
        raise Framework.FrameworkError('install_into_reference: the reference Implementations module exposes no Framework module')
    Framework.bind_host_framework(reference_framework_module)
    reference_implementations_module.Methods.modules['NeRF'] = Methods.import_method('NeRF')


def uninstall_from_reference(reference_implementations_module) -> None:
    """Undo install_into_reference (tests)."""
    reference_implementations_module.Methods.modules.pop('NeRF', None)
    Framework.unbind_host_framework()
