"""Method plugin triple, as resolved by the reference's ``Implementations.Methods.import_method('NeRF')``
(src/Methods/NeRF/__init__.py:5-7)."""
from .Model import NeRF
from .Renderer import NeRFRenderer
from .Trainer import NeRFTrainer

MODEL = NeRF
RENDERER = NeRFRenderer
TRAINING_INSTANCE = NeRFTrainer
