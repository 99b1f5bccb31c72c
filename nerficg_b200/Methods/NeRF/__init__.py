"""The plugin triple that ``Implementations.Methods.import_method('NeRF')`` looks up by attribute name
(reference src/Methods/NeRF/__init__.py:5-7): MODEL / RENDERER / TRAINING_INSTANCE."""
from . import Model as _model, Renderer as _renderer, Trainer as _trainer

MODEL, RENDERER, TRAINING_INSTANCE = _model.NeRF, _renderer.NeRFRenderer, _trainer.NeRFTrainer
__all__ = ['MODEL', 'RENDERER', 'TRAINING_INSTANCE']
