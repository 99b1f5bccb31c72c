"""Pseudo-colouring of the test-set output path (reference src/Visual/utils.py:8-34 ``apply_color_map`` and
src/Visual/ColorMap.py ``ColorMap.apply``).  Only what ``NeRFRenderer.postprocess_outputs`` uses:
the SPECTRAL map (and Grayscale), nearest-entry lookup, optional mask.

The reference ships SPECTRAL as a 256-entry table; it is matplotlib's ``Spectral``: the eleven ColorBrewer "Spectral"
anchor colours interpolated linearly over 256 entries (checked against the reference's table: max difference 6e-8), so the
table is generated here instead of stored."""
from __future__ import annotations

import torch

from .. import Framework

# ColorBrewer 11-class "Spectral" (8-bit sRGB)
_SPECTRAL_ANCHORS = ((158, 1, 66), (213, 62, 79), (244, 109, 67), (253, 174, 97), (254, 224, 139), (255, 255, 191),
                     (230, 245, 152), (171, 221, 164), (102, 194, 165), (50, 136, 189), (94, 79, 162))
_luts: dict[tuple[str, str], torch.Tensor] = {}


def spectral_lut(device: torch.device | str = 'cpu') -> torch.Tensor:
    """(256, 3) fp32 table: entry i is the piecewise-linear interpolation of the anchors at i / 255."""
    key = ('SPECTRAL', str(device))
    if key not in _luts:
        anchors = torch.tensor(_SPECTRAL_ANCHORS, dtype=torch.float64) / 255.0
        x = torch.linspace(0.0, 1.0, 256, dtype=torch.float64) * (len(_SPECTRAL_ANCHORS) - 1)
        lo = x.floor().clamp(max=len(_SPECTRAL_ANCHORS) - 2).long()
        frac = (x - lo).unsqueeze(1)
        _luts[key] = (anchors[lo] * (1.0 - frac) + anchors[lo + 1] * frac).to(dtype=torch.float32, device=device)
    return _luts[key]


@torch.no_grad()
def apply_color_map(color_map: str, image: torch.Tensor, min_max: tuple[float, float] | None = None,
                    mask: torch.Tensor | None = None, invert: bool = False) -> torch.Tensor:
    """(1, H, W) -> (3, H, W) in [0, 1].  ``min_max`` None normalises to the range of the (masked, mask > 0.99) image."""
    if image.ndim != 3 or image.shape[0] != 1:
        raise Framework.RendererError('apply_color_map: input must have shape (1, H, W)')
    if min_max is None:
        masked = image[mask > 0.99] if mask is not None else image
        min_val = masked.min() if masked.numel() > 0 else 0.0
        max_val = masked.max() if masked.numel() > 0 else 1.0
    else:
        min_val, max_val = min_max
    image = torch.clamp((image - min_val) / (max_val - min_val), min=0.0, max=1.0)
    if invert:
        image = 1.0 - image
    if color_map == 'Grayscale':
        out = image.expand((3, image.shape[1], image.shape[2])).clone()
    elif color_map == 'SPECTRAL':
        index = (image * 255).int().flatten()   # truncation, like the reference's index_select
        out = torch.index_select(spectral_lut(image.device), 0, index).reshape(*image.shape[1:], 3).permute(2, 0, 1).contiguous()
    else:
        raise Framework.RendererError(f'apply_color_map: unsupported colour map "{color_map}" (SPECTRAL, Grayscale)')
    if mask is not None:
        out *= mask
    return out
