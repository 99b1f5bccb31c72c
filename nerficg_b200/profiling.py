"""NVTX ranges per stage of the hot path (profiling row): K0..K8 show up as named ranges in an
Nsight Systems / Nsight Compute timeline.  Off by default (the captured training step issues no host work at replay, and the
eager paths should not pay for string marshalling); switch on with NERF_B200_NVTX=1 or ``profiling.enable()``."""
from __future__ import annotations

import contextlib
import os

import torch

_enabled = os.environ.get('NERF_B200_NVTX', '0') == '1'


def enable(on: bool = True) -> None:
    global _enabled
    _enabled = bool(on)


def enabled() -> bool:
    return _enabled


@contextlib.contextmanager
def stage(name: str):
    """``with profiling.stage('K3 mlp_fwd fine'):`` -- an NVTX range around the launches of one stage."""
    if _enabled and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield
