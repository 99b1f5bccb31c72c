"""Generates tests/golden/postprocess.pt by running the UNMODIFIED reference's ``apply_color_map`` (src/Visual/utils.py:8-34,
src/Visual/ColorMap.py) on CPU -- the depth pseudo-colouring of ``NeRFRenderer.postprocess_outputs``
(src/Methods/NeRF/Renderer.py:142-165).  TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference).

    python oracle/make_golden_postprocess.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.ref_loader import load_reference  # noqa: E402


def main() -> None:
    load_reference()
    from Visual.utils import apply_color_map           # the reference's module (src/ is on sys.path now)
    g = torch.Generator().manual_seed(5)
    h, w = 13, 17
    depth = 1.5 + 5.0 * torch.rand(1, h, w, generator=g)           # partly outside [near, far] = [2, 6]: exercises the clamp
    depth[0, 0, :4] = torch.tensor([2.0, 6.0, 4.0, 0.0])
    alpha = torch.rand(1, h, w, generator=g)
    alpha[0, 1, :3] = torch.tensor([0.0, 1.0, 0.995])
    out = {
        'depth': depth, 'alpha': alpha, 'near': 2.0, 'far': 6.0,
        'spectral_masked': apply_color_map(color_map='SPECTRAL', image=depth.clone(), min_max=(2.0, 6.0), mask=alpha.clone()),
        'spectral_auto': apply_color_map(color_map='SPECTRAL', image=depth.clone(), min_max=None, mask=alpha.clone()),
        'spectral_plain': apply_color_map(color_map='SPECTRAL', image=depth.clone(), min_max=(2.0, 6.0)),
        'gray_inverted': apply_color_map(color_map='Grayscale', image=depth.clone(), min_max=(2.0, 6.0), invert=True),
    }
    path = ROOT / 'tests' / 'golden' / 'postprocess.pt'
    torch.save(out, path)
    print('wrote', path, path.stat().st_size, 'bytes')


if __name__ == '__main__':
    main()
