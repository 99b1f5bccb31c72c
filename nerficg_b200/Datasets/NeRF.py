"""Blender-synthetic ("nerf_synthetic") scenes on disk (reference src/Datasets/NeRF.py:45-107).

``transforms_{train,val,test}.json`` give ``camera_angle_x`` and one 4x4 ``transform_matrix`` per frame (Blender world,
OpenGL camera); images are RGBA PNGs ``<file_path>.png``; test frames also have ``<file_path>_depth_*.png``.
Poses are converted like the reference: c2w = B2C @ transform_matrix @ GL2C^T with B2C the Blender -> Colmap world
permutation and GL2C = diag(1, -1, -1, 1); focal = 0.5 / tan(camera_angle_x / 2) * width.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch
from torchvision import io

from .. import Framework
from ..Cameras.Perspective import PerspectiveCamera, fov_to_focal
from .Base import BaseDataset
from .utils import View

_CAM_TRANSFORM = np.diag([1.0, -1.0, -1.0, 1.0])                       # OpenGL -> Colmap camera axes
_WORLD_TRANSFORM = np.array([[1.0, 0.0, 0.0, 0.0],
                             [0.0, 0.0, -1.0, 0.0],
                             [0.0, 1.0, 0.0, 0.0],
                             [0.0, 0.0, 0.0, 1.0]])                     # Blender -> Colmap world axes


def load_image(path: Path) -> torch.Tensor:
    """C x H x W float image in [0, 1] (8- or 16-bit PNG), reference Datasets/utils.py:66-74."""
    try:
        image = io.decode_image(str(path), mode=io.ImageReadMode.UNCHANGED)
    except Exception:
        raise Framework.DatasetError(f'Failed to load image file: "{path}"')
    return image.float() / (65535 if image.dtype == torch.uint16 else 255)


def load_nerf_depth(path: Path) -> torch.Tensor:
    """Depth map of a test frame: the original NeRF Blender files store 1 - depth / 8 in the first channel
    (reference NeRF.py:23-33)."""
    return -(load_image(path)[:1] - 1.0) * 8.0


@Framework.Configurable.configure(
    PATH='dataset/nerf_synthetic/lego',
    NORMALIZE_CUBE=4.0 / 1.5,
    NEAR_PLANE=2.0,
    FAR_PLANE=6.0,
)
class CustomDataset(BaseDataset):
    def load(self):
        camera = None
        self.bounding_box = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]], dtype=torch.float32)
        data: dict[str, list[View]] = {subset: [] for subset in self.subsets}
        for subset in self.subsets:
            meta_path = self.dataset_path / f'transforms_{subset}.json'
            try:
                with open(meta_path) as f:
                    meta = json.load(f)
            except IOError:
                raise Framework.DatasetError(f'Invalid dataset metadata file path "{meta_path}"')
            for frame_idx, frame in enumerate(meta['frames']):
                rgba_path = self.dataset_path / f'{frame["file_path"]}.png'
                rgba = load_image(rgba_path)
                if rgba.shape[0] != 4:
                    raise Framework.DatasetError(f'"{rgba_path}" is not an RGBA image')
                height, width = rgba.shape[1:]
                focal = fov_to_focal(float(meta['camera_angle_x'])) * width
                if camera is None:
                    camera = PerspectiveCamera(shared_settings=self._camera_settings, width=width, height=height, focal_x=focal, focal_y=focal)
                elif camera.focal_x != focal or camera.width != width or camera.height != height:
                    raise Framework.DatasetError('The NeRF loader requires all views to have the same image size and focal length.')
                c2w = _WORLD_TRANSFORM @ np.asarray(frame['transform_matrix'], dtype=np.float64) @ _CAM_TRANSFORM.T
                depth = None
                if subset == 'test':  # the synthetic NeRF dataset ships depth for the test set only
                    depth_path = next(self.dataset_path.glob(f'{frame["file_path"]}_depth_*.png'), None)
                    if depth_path is None:
                        raise Framework.DatasetError(f'no depth map "{frame["file_path"]}_depth_*.png" for test frame {frame_idx}')
                    depth = load_nerf_depth(depth_path)
                data[subset].append(View(camera, c2w, rgb=rgba[:3].contiguous(), alpha=rgba[3:4].contiguous(), depth=depth,
                                         frame_idx=frame_idx))
        if camera is None:
            raise Framework.DatasetError(f'no frames found under "{self.dataset_path}"')
        return [camera], data
