"""Flat fp32 parameter buffers of NeRFBlocks, in the layout the CUDA kernels read (nerf_param_layout).

The 24 tensors of one block keep the reference's names, shapes and row-major (out, in) layout
(reference src/Methods/NeRF/Model.py:35-54); they are views into one buffer so the
kernels take a single pointer, Adam and NCCL see one contiguous gradient, and checkpoints stay
key-compatible with nerficg.
"""
from __future__ import annotations

import torch

from . import _lib

# (state-dict key suffix, shape) in torch registration order of the reference NeRFBlock
TENSOR_SPECS: list[tuple[str, tuple[int, ...]]] = []
for _l in range(8):
    _in = 63 if _l == 0 else (319 if _l == 5 else 256)
    TENSOR_SPECS += [(f'initial_layers.{_l}.0.weight', (256, _in)), (f'initial_layers.{_l}.0.bias', (256,))]
TENSOR_SPECS += [('feature_layer.weight', (256, 256)), ('feature_layer.bias', (256,)),
                 ('density_layer.weight', (1, 256)), ('density_layer.bias', (1,)),
                 ('color_layers.0.weight', (128, 283)), ('color_layers.0.bias', (128,)),
                 ('color_layers.2.weight', (3, 128)), ('color_layers.2.bias', (3,))]

_layout = None


def layout() -> tuple[list[int], list[int], int]:
    """(offsets, sizes, total) in floats, queried from the C library (single source of truth)."""
    global _layout
    if _layout is None:
        off, size, total = _lib.param_layout()
        for (name, shape), n in zip(TENSOR_SPECS, size):
            assert int(torch.Size(shape).numel()) == n, f'layout mismatch for {name}'
        _layout = (off, size, total)
    return _layout


def views(flat: torch.Tensor) -> dict[str, torch.Tensor]:
    """Named views (reference key suffixes) into a flat block buffer."""
    off, size, total = layout()
    assert flat.numel() == total and flat.is_contiguous()
    return {name: flat[o:o + n].view(shape) for (name, shape), o, n in zip(TENSOR_SPECS, off, size)}


def flatten_state_dict(sd: dict, prefix: str, device=None) -> torch.Tensor:
    """Copies one block of a nerficg state dict ('nerf.' / 'coarse_nerf.') into a new flat buffer."""
    _, _, total = layout()
    flat = torch.zeros(total, dtype=torch.float32, device=device)
    for name, view in views(flat).items():
        view.copy_(sd[prefix + name])
    return flat
