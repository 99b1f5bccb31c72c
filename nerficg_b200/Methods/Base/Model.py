"""Base model: metadata + checkpoint format of the reference (src/Methods/Base/Model.py:14-111).

Checkpoints are ``torch.save`` dicts with ``model_state_dict`` plus ``model_name, creation_date,
num_iterations_trained, output_directory`` and every MODEL config key, so files written by nerficg load
here and vice versa."""
from __future__ import annotations

import datetime
from abc import ABC, abstractmethod
from pathlib import Path

import torch

from ... import Framework
from ...Logging import Logger

_META = ['model_name', 'creation_date', 'num_iterations_trained', 'output_directory']


class BaseModel(Framework.Configurable, ABC, torch.nn.Module):
    def __init__(self, name: str = None) -> None:
        Framework.Configurable.__init__(self, 'MODEL')
        ABC.__init__(self)
        torch.nn.Module.__init__(self)
        self.model_name: str = name if name is not None else 'Default'
        self.creation_date: str = f'{datetime.datetime.now():%Y-%m-%d-%H-%M-%S}'
        self.num_iterations_trained: int = 0
        self.output_directory: Path = (Framework.Directories.OUTPUT_DIR / str(Framework.config.GLOBAL.METHOD_TYPE)
                                       / f'{self.model_name}_{self.creation_date}')

    @abstractmethod
    def build(self) -> 'BaseModel':
        return self

    def forward(self) -> None:
        Logger.log_error('Model cannot be executed directly. Use a Renderer instead.')

    @classmethod
    def load(cls, checkpoint_name: str | Path | None, map_location='cpu') -> 'BaseModel':
        if checkpoint_name is None or str(checkpoint_name).split('.')[-1] != 'pt':
            raise Framework.ModelError(f'Invalid model checkpoint: "{checkpoint_name}"')
        try:
            path = Path(checkpoint_name)
            if not path.is_absolute():
                path = Framework.Directories.NERFICG_ROOT / path
            checkpoint = torch.load(path, map_location=map_location, weights_only=False)
        except IOError as e:
            raise Framework.ModelError(f'failed to load model from file: "{e}"')
        return cls.from_checkpoint_dict(checkpoint)

    @classmethod
    def from_checkpoint_dict(cls, checkpoint: dict) -> 'BaseModel':
        """Rebuilds a model from the dict ``save`` writes -- the reference's checkpoint format (Base/Model.py:60-111)."""
        model = cls()
        for key in _META + list(cls.get_default_parameters().keys()):
            if key in checkpoint:
                model.__dict__[key] = checkpoint[key]
            else:
                Logger.log_warning(f'failed to load model parameter "{key}" -> using default value "{model.__dict__[key]}"')
        model.build()
        missing, unexpected = model.load_state_dict(checkpoint['model_state_dict'], strict=False)
        for key in missing:
            Logger.log_warning(f'missing key in model checkpoint: "{key}"')
        for key in unexpected:
            Logger.log_warning(f'unexpected key in model checkpoint: "{key}"')
        device = Framework.config.GLOBAL.get('DEFAULT_DEVICE')
        return model.to(device) if device is not None else model

    def checkpoint_dict(self) -> dict:
        checkpoint = {'model_state_dict': self.state_dict()}
        for key in _META + list(type(self).get_default_parameters().keys()):
            checkpoint[key] = self.__dict__[key]
        return checkpoint

    def save(self, path: Path) -> None:
        try:
            torch.save(self.checkpoint_dict(), path)
        except IOError as e:
            Logger.log_warning(f'failed to save model: "{e}"')
