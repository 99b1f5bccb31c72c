"""Generates tests/golden/rays.pt by running the UNMODIFIED reference's ray generation on CPU.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_rays.py        (build container only: needs /root/reference)

Reference code exercised: ``PerspectiveCamera.compute_local_ray_directions`` (src/Cameras/Perspective.py:64-94) and
``View.get_rays`` / ``View.cam_to_world`` (src/Datasets/utils.py:1033-1074).  Two cameras: the Lego-shaped 100x100 one
(a random subset of its pixels is stored) and a small 37x23 one with fx != fy and an off-centre principal point
(stored completely).
"""
from __future__ import annotations

import math
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = ROOT / 'tests' / 'golden'


def look_at_pose(theta: float, phi: float, radius: float = 4.0311) -> np.ndarray:
    pos = radius * np.array([math.cos(phi) * math.cos(theta), math.cos(phi) * math.sin(theta), math.sin(phi)])
    fwd = -pos / np.linalg.norm(pos)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4, dtype=np.float64)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, pos
    return c2w


def main() -> None:
    ref = load_reference()
    DS, Camera, Shared = ref['ds_utils'], ref['PerspectiveCamera'], ref['SharedCameraSettings']
    shared = Shared(background_color=torch.ones(3), near_plane=2.0, far_plane=6.0)
    cases = []
    g = torch.Generator().manual_seed(0)
    for (w, h, fx, fy, cx, cy, theta, phi, keep) in (
            (100, 100, 0.5 / math.tan(0.5 * 0.6911112070083618) * 100, None, None, None, 0.9, 0.6, 1500),
            (37, 23, 41.7, 38.2, 17.3, 12.9, -2.1, 0.2, None)):
        cam = Camera(shared_settings=shared, width=w, height=h, focal_x=fx, focal_y=fy, center_x=cx, center_y=cy)
        c2w = look_at_pose(theta, phi)
        view = DS.View(camera=cam, camera_index=0, frame_idx=0, global_frame_idx=0, c2w=c2w)
        rays = view.get_rays()
        ids = torch.arange(w * h) if keep is None else torch.randint(0, w * h, (keep,), generator=g)
        cases.append({'width': w, 'height': h, 'focal_x': cam.focal_x, 'focal_y': cam.focal_y, 'center_x': cam.center_x,
                      'center_y': cam.center_y, 'c2w': torch.from_numpy(c2w), 'pixel_ids': ids,
                      'origin': rays.origin[ids].clone(), 'direction': rays.direction[ids].clone(),
                      'view_direction': rays.view_direction[ids].clone()})
        print(w, h, rays.direction.shape, rays.direction[0], rays.view_direction[0])
    torch.save(cases, OUT / 'rays.pt')
    print('wrote', OUT / 'rays.pt', (OUT / 'rays.pt').stat().st_size, 'bytes')


if __name__ == '__main__':
    main()
