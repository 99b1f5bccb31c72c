"""A tiny Blender-synthetic-format scene for the loader tests (and for oracle/make_golden_loader.py, which runs the
reference's own loader on it)."""
import json
import math
from pathlib import Path

import numpy as np
import torch
from PIL import Image

ANGLE_X = 0.6911112070083618


def write_scene(root: Path) -> None:
    """2 train, 1 val, 2 test frames of 8 x 6 RGBA PNGs with Blender-style poses; test frames get depth PNGs."""
    g = torch.Generator().manual_seed(42)
    for subset, count in (('train', 2), ('val', 1), ('test', 2)):
        (root / subset).mkdir(parents=True, exist_ok=True)
        frames = []
        for i in range(count):
            th, ph = 0.7 * i + 0.3 * len(subset), 0.4 + 0.1 * i
            pos = 4.0311 * np.array([math.cos(ph) * math.cos(th), math.cos(ph) * math.sin(th), math.sin(ph)])
            back = pos / np.linalg.norm(pos)                     # OpenGL camera looks down -z
            right = np.cross([0.0, 0.0, 1.0], back)
            right /= np.linalg.norm(right)
            up = np.cross(back, right)
            m = np.eye(4)
            m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, back, pos
            frames.append({'file_path': f'./{subset}/r_{i}', 'rotation': 0.1, 'transform_matrix': m.tolist()})
            Image.fromarray(torch.randint(0, 256, (6, 8, 4), generator=g, dtype=torch.uint8).numpy(), 'RGBA').save(root / subset / f'r_{i}.png')
            if subset == 'test':
                Image.fromarray(torch.randint(0, 256, (6, 8, 4), generator=g, dtype=torch.uint8).numpy(), 'RGBA').save(root / subset / f'r_{i}_depth_0001.png')
        (root / f'transforms_{subset}.json').write_text(json.dumps({'camera_angle_x': ANGLE_X, 'frames': frames}))


def summarize(dataset) -> dict:
    out = {}
    for subset in ('train', 'val', 'test'):
        views = dataset.data[subset]
        out[subset] = {
            'c2w': torch.stack([torch.as_tensor(np.asarray(v.c2w_numpy if hasattr(v, 'c2w_numpy') else v.c2w), dtype=torch.float64) for v in views]),
            'rgb': torch.stack([v.rgb.cpu() for v in views]), 'alpha': torch.stack([v.alpha.cpu() for v in views]),
            'depth': torch.stack([v.depth.cpu() for v in views]) if subset == 'test' else None}
    cam = dataset.default_camera
    out['camera'] = (cam.width, cam.height, cam.focal_x, cam.focal_y, cam.center_x, cam.center_y, cam.near_plane, cam.far_plane,
                     cam.background_color.cpu().tolist())
    return out
