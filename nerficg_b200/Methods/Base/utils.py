"""Training-callback decorators (reference src/Methods/Base/utils.py:36-92): methods of a trainer are tagged
with a type (-1 before, 0 during, 1 after training), a priority (higher first), and optional start / end /
stride in iterations; string values name trainer attributes that are resolved when training starts."""
from __future__ import annotations

from typing import Callable


def _callback(callback_type: int, active=True, priority: int = 50, start_iteration=None, end_iteration=None,
              iteration_stride=None) -> Callable:
    def decorator(function: Callable) -> Callable:
        def wrapper(*args, **kwargs):
            return function(*args, **kwargs)
        wrapper.callback_type = callback_type
        wrapper.active = active
        wrapper.priority = priority
        wrapper.start_iteration = start_iteration
        wrapper.end_iteration = end_iteration
        wrapper.iteration_stride = iteration_stride
        wrapper.__name__ = function.__name__
        return wrapper
    return decorator


def training_callback(active=True, priority: int = 50, start_iteration=None, end_iteration=None, iteration_stride=None):
    return _callback(0, active, priority, start_iteration, end_iteration, iteration_stride)


def pre_training_callback(active=True, priority: int = 50):
    return _callback(-1, active, priority)


def post_training_callback(active=True, priority: int = 50):
    return _callback(1, active, priority)
