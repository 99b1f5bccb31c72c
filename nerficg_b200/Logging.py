"""Level-based console logger with the reference's interface (reference src/Logging.py:8-51):
``Logger.set_mode(level)`` and ``log / log_info / log_warning / log_error / log_debug / log_progress``."""
from __future__ import annotations

import sys


class Logger:
    MODE_SILENT, MODE_NORMAL, MODE_VERBOSE, MODE_DEBUG = range(4)
    _level = MODE_NORMAL

    @classmethod
    def set_mode(cls, lvl: int) -> None:
        cls._level = lvl if lvl in range(4) else cls.MODE_NORMAL

    @classmethod
    def _emit(cls, min_level: int, tag: str, msg) -> None:
        if cls._level >= min_level:
            print(f'{tag}{msg}', file=sys.stdout, flush=True)

    @classmethod
    def log(cls, msg) -> None:
        cls._emit(cls.MODE_NORMAL, '', msg)

    @classmethod
    def log_error(cls, msg) -> None:
        cls._emit(cls.MODE_VERBOSE, 'ERROR: ', msg)

    @classmethod
    def log_info(cls, msg) -> None:
        cls._emit(cls.MODE_VERBOSE, 'INFO: ', msg)

    @classmethod
    def log_warning(cls, msg) -> None:
        cls._emit(cls.MODE_VERBOSE, 'WARNING: ', msg)

    @classmethod
    def log_debug(cls, msg) -> None:
        cls._emit(cls.MODE_DEBUG, 'DEBUG: ', msg)

    @classmethod
    def log_progress(cls, iterable, **kwargs):
        if cls._level >= cls.MODE_NORMAL:
            try:
                from tqdm.auto import tqdm
                return tqdm(iterable, file=sys.stdout, dynamic_ncols=True, **kwargs)
            except ImportError:
                pass
        return iterable
