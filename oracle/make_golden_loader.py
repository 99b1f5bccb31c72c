"""Generates tests/golden/blender_loader.pt by running the UNMODIFIED reference's Blender-synthetic loader
(src/Datasets/NeRF.py:45-107, src/Datasets/Base.py) on the tiny scene of tests/blender_scene.py.  TEST INFRASTRUCTURE ONLY;
build container only (needs /root/reference).

    python oracle/make_golden_loader.py
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
from blender_scene import summarize, write_scene  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402


def main() -> None:
    ref = load_reference()
    with tempfile.TemporaryDirectory() as tmp:
        write_scene(Path(tmp))
        ref['Framework'].config.DATASET.PATH = tmp
        ref['Framework'].config.DATASET.NORMALIZE_CUBE = None
        from Datasets.NeRF import CustomDataset          # the reference's module (src/ is on sys.path now)
        out = summarize(CustomDataset(tmp))
    path = ROOT / 'tests' / 'golden' / 'blender_loader.pt'
    torch.save(out, path)
    print('wrote', path, path.stat().st_size, 'bytes')


if __name__ == '__main__':
    main()
