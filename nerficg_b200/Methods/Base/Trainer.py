"""Base trainer: callback-driven loop of the reference (src/Methods/Base/Trainer.py:225-291) without its
wandb / GUI / timing side channels (out of scope)."""
from __future__ import annotations

import inspect
from operator import attrgetter
from pathlib import Path
from typing import Callable

import torch

from ... import Framework
from ...Logging import Logger
from .Model import BaseModel
from .Renderer import BaseRenderer
from .utils import post_training_callback


@Framework.Configurable.configure(
    LOAD_CHECKPOINT=None,
    MODEL_NAME='Default',
    NUM_ITERATIONS=1,
    RUN_VALIDATION=False,
    BACKUP=Framework.ConfigParameterList(FINAL_CHECKPOINT=True, RENDER_TESTSET=True, RENDER_TRAINSET=False, RENDER_VALSET=False,
                                         INTERMEDIATE_RENDERINGS=True, VISUALIZE_ERRORS=False, INTERVAL=-1, TRAINING_STATE=False),
    WANDB=Framework.ConfigParameterList(ACTIVATE=False),
)
class BaseTrainer(Framework.Configurable, torch.nn.Module):
    def __init__(self, model: BaseModel, renderer: BaseRenderer, write_outputs: bool = False) -> None:
        Framework.Configurable.__init__(self, 'TRAINING')
        torch.nn.Module.__init__(self)
        self.model = model
        self.renderer = renderer
        self.write_outputs = write_outputs
        self.output_directory: Path = model.output_directory
        self.checkpoint_directory: Path = self.output_directory / 'checkpoints'
        if write_outputs:
            self.checkpoint_directory.mkdir(parents=True, exist_ok=True)

    # ---- '.train' checkpoints (reference Base/Trainer.py:94-111) -------------------------------------------
    # The reference pickles the whole trainer object.  Ours owns CUDA graphs and ctypes handles, so the file holds the
    # state needed to resume instead: model checkpoint dict, optimiser state_dict (torch.optim.Adam layout), schedule
    # position and the TRAINING parameters; ``load`` rebuilds model -> renderer -> trainer from it.
    @classmethod
    def _method_classes(cls) -> tuple[type, type]:
        raise Framework.CheckpointError(f'{cls.__name__} does not define its MODEL / RENDERER classes')

    def save(self, path: Path) -> None:
        try:
            state = {'format': 'nerficg_b200.train/1', 'trainer_class': type(self).__name__,
                     'model': self.model.checkpoint_dict(),
                     'optimizer': self.optimizer.state_dict() if hasattr(self, 'optimizer') else None,
                     'training_parameters': {k: (v.toDict() if isinstance(v, Framework.ConfigParameterList) else v)
                                             for k, v in self.__dict__.items() if k in type(self).get_default_parameters()}}
            torch.save(state, path)
        except IOError as e:
            raise Framework.CheckpointError(f'Failed to save checkpoint "{e}"')

    @classmethod
    def load(cls, checkpoint_name: str | Path) -> 'BaseTrainer':
        if checkpoint_name is None or str(checkpoint_name).split('.')[-1] != 'train':
            raise Framework.CheckpointError(f'Invalid checkpoint name "{checkpoint_name}"')
        try:
            path = Path(checkpoint_name)
            if not path.is_absolute():
                path = Framework.Directories.NERFICG_ROOT / path
            state = torch.load(path, map_location='cpu', weights_only=False)
        except IOError as e:
            raise Framework.CheckpointError(f'Failed to load checkpoint "{e}"')
        if not isinstance(state, dict) or state.get('format') != 'nerficg_b200.train/1':
            raise Framework.CheckpointError(f'"{checkpoint_name}" is not a nerficg_b200 training checkpoint (a pickled reference '
                                            'trainer cannot be resumed on the CUDA path; load its model checkpoint instead)')
        model_cls, renderer_cls = cls._method_classes()
        model = model_cls.from_checkpoint_dict(state['model'])
        trainer = cls(model=model, renderer=renderer_cls(model))
        for key, value in state['training_parameters'].items():   # the run that is resumed keeps its own hyper-parameters
            trainer.__dict__[key] = Framework.ConfigParameterList.fromDict(value) if isinstance(value, dict) else value
        if state.get('optimizer') is not None and hasattr(trainer, 'optimizer'):
            trainer.optimizer.load_state_dict(state['optimizer'])
        return trainer

    def _callbacks(self, callback_type: int) -> list[Callable]:
        found = []
        for _, fn in inspect.getmembers(type(self), predicate=inspect.isfunction):
            if getattr(fn, 'callback_type', None) != callback_type:
                continue
            for attr in ('active', 'start_iteration', 'end_iteration', 'iteration_stride'):
                value = getattr(fn, attr)
                if isinstance(value, str):
                    try:
                        setattr(fn, attr, attrgetter(value)(self))
                    except AttributeError:
                        raise Framework.TrainingError(f'callback "{fn.__name__}" refers to unknown config parameter "{value}"')
            if fn.iteration_stride is not None and fn.iteration_stride <= 0:
                continue
            if fn.active:
                found.append(fn)
        return sorted(found, key=lambda c: c.priority, reverse=True)

    def run(self, dataset) -> None:
        Logger.log(f'starting training for model: {self.model.model_name}')
        start = iteration = self.model.num_iterations_trained
        if start <= 0:
            for cb in self._callbacks(-1):
                cb(self, start, dataset)
        try:
            training = self._callbacks(0)
            for iteration in Logger.log_progress(range(start, self.NUM_ITERATIONS), desc='training', miniters=10):
                for cb in training:
                    if ((cb.start_iteration is not None and iteration < cb.start_iteration) or
                            (cb.end_iteration is not None and iteration > cb.end_iteration) or
                            (cb.iteration_stride is not None and (iteration - (cb.start_iteration or 0)) % cb.iteration_stride != 0)):
                        continue
                    cb(self, iteration, dataset)
                self.model.num_iterations_trained += 1
        except KeyboardInterrupt:
            Logger.log_warning('training manually interrupted')
        for cb in self._callbacks(1):
            cb(self, iteration + 1, dataset)
        Logger.log('training finished successfully')

    @post_training_callback(active='BACKUP.FINAL_CHECKPOINT', priority=100)
    def _save_final_checkpoint(self, _, dataset) -> None:
        if self.write_outputs:
            self.model.save(self.checkpoint_directory / 'final.pt')
