"""CPU: the Blender-synthetic loader (nerficg_b200/Datasets/NeRF.py, reference src/Datasets/NeRF.py:45-107) on a tiny scene
written to a temporary directory; poses, intrinsics and decoded images against a committed golden produced by the reference's own loader on the same
scene (tests/golden/blender_loader.pt, written by oracle/make_golden_loader.py)."""
import math
from pathlib import Path

import pytest
import torch

from blender_scene import ANGLE_X, summarize, write_scene
from nerficg_b200 import Framework

GOLDEN = Path(__file__).resolve().parent / 'golden' / 'blender_loader.pt'

def load_ours(root: Path, **dataset_overrides):
    Framework.load_config(None, {'GLOBAL.LOG_LEVEL': 0, 'DATASET.PATH': str(root), 'DATASET.BACKGROUND_COLOR': [1.0, 1.0, 1.0],
                                 'DATASET.NORMALIZE_CUBE': None, **{f'DATASET.{k}': v for k, v in dataset_overrides.items()}})
    from nerficg_b200.Datasets.NeRF import CustomDataset
    return CustomDataset(str(root))


def test_loader_matches_reference_golden(tmp_path):
    write_scene(tmp_path)
    gold = torch.load(GOLDEN, weights_only=False)
    ours = summarize(load_ours(tmp_path))
    assert ours['camera'][:2] == gold['camera'][:2] and ours['camera'][8] == gold['camera'][8]
    assert ours['camera'][2:8] == pytest.approx(gold['camera'][2:8], rel=1e-12)
    for subset in ('train', 'val', 'test'):
        assert (ours[subset]['c2w'] - gold[subset]['c2w']).abs().max() <= 1e-6      # our View keeps c2w in float32
        assert torch.equal(ours[subset]['rgb'], gold[subset]['rgb']) and torch.equal(ours[subset]['alpha'], gold[subset]['alpha'])
    assert torch.equal(ours['test']['depth'], gold['test']['depth'])


def test_loader_interface_and_errors(tmp_path):
    write_scene(tmp_path)
    ds = load_ours(tmp_path)
    assert (len(ds.train()), len(ds.eval()), len(ds.test())) == (2, 1, 2)
    assert ds.train().get_total_ray_count() == 2 * 8 * 6
    view = ds.test()[1]
    assert view.rgb.shape == (3, 6, 8) and view.alpha.shape == (1, 6, 8) and view.depth.shape == (1, 6, 8)
    assert ds.default_camera.focal_x == pytest.approx(0.5 / math.tan(0.5 * ANGLE_X) * 8)
    # camera looks at the origin along +z of the converted pose, y is down
    c2w = view.c2w.double()
    assert torch.allclose(torch.nn.functional.normalize(-c2w[:3, 3], dim=0), c2w[:3, 2], atol=1e-6)
    with pytest.raises(Framework.DatasetError):
        ds.set_mode('nope')
    from nerficg_b200.Implementations import Datasets
    assert type(Datasets.get_dataset('NeRF', str(tmp_path))) is type(ds)
    with pytest.raises(Framework.DatasetError):
        Datasets.get_dataset_class('Colmap')
    # pose normalisation (reference Base.py:218-244): camera positions fit a cube of the requested side, planes scale along
    ds2 = load_ours(tmp_path, NORMALIZE_CUBE=2.0, NORMALIZE_RECENTER=True)
    pos = torch.stack([v.position for v in ds2.data['train']])
    assert float((pos.max(0).values - pos.min(0).values).max()) == pytest.approx(2.0, rel=1e-5)
    assert torch.allclose(pos.max(0).values + pos.min(0).values, torch.zeros(3), atol=1e-5)
    assert ds2.default_camera.near_plane / ds.default_camera.near_plane == pytest.approx(ds2.default_camera.far_plane / ds.default_camera.far_plane)
    (tmp_path / 'test' / 'r_1_depth_0001.png').unlink()
    with pytest.raises(Framework.DatasetError):
        load_ours(tmp_path)
    with pytest.raises(Framework.DatasetError):
        load_ours(tmp_path / 'missing')
