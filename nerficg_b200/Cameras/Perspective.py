"""Pinhole camera of the NeRF path (reference src/Cameras/{Base,Perspective,utils}.py, distortion-free subset).

Lego-shaped data has no lens distortion, so only the ideal pinhole model is kept; ray generation
follows ``compute_local_ray_directions`` (Perspective.py:64-94): pixel-centre rays with camera-space
z = 1 that are NOT normalised.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from .. import Framework


def fov_to_focal(fov: float, degrees: bool = False) -> float:
    """Field of view -> normalised focal length (reference Cameras/utils.py:231-234)."""
    return 0.5 / math.tan(0.5 * (math.radians(fov) if degrees else fov))


def focal_to_fov(focal: float, degrees: bool = False) -> float:
    fov = 2 * math.atan(0.5 / focal)
    return math.degrees(fov) if degrees else fov


@dataclass
class SharedCameraSettings:
    """Background colour and clipping planes shared by all cameras of a dataset (Cameras/utils.py:163-178)."""
    background_color: torch.Tensor
    near_plane: float
    far_plane: float

    def __post_init__(self):
        if tuple(self.background_color.shape) != (3,):
            raise Framework.CameraError(f'background_color must have shape (3,), got {tuple(self.background_color.shape)}')
        if self.near_plane <= 0 or self.far_plane <= self.near_plane:
            raise Framework.CameraError('invalid clipping planes: need 0 < near_plane < far_plane')


@dataclass(kw_only=True)
class PerspectiveCamera:
    shared_settings: SharedCameraSettings
    width: int
    height: int
    focal_x: float = None
    focal_y: float = None
    center_x: float = None
    center_y: float = None
    _local_ray_directions_cache: object = field(init=False, default=None, repr=False)

    def __post_init__(self) -> None:
        if self.focal_x is None and self.focal_y is None:
            self.focal_x = self.focal_y = fov_to_focal(45.0, degrees=True) * self.height
        elif self.focal_x is None:
            self.focal_x = self.focal_y
        elif self.focal_y is None:
            self.focal_y = self.focal_x
        if self.center_x is None:
            self.center_x = self.width / 2
        if self.center_y is None:
            self.center_y = self.height / 2

    @property
    def background_color(self) -> torch.Tensor:
        return self.shared_settings.background_color

    @background_color.setter
    def background_color(self, color: torch.Tensor) -> None:
        self.shared_settings.background_color = color.to(self.shared_settings.background_color)

    @property
    def near_plane(self) -> float:
        return self.shared_settings.near_plane

    @property
    def far_plane(self) -> float:
        return self.shared_settings.far_plane

    def compute_local_ray_directions(self, through_pixel_center: bool = True, device=None) -> torch.Tensor:
        """(H*W, 3) camera-space directions ((x+.5-cx)/fx, (y+.5-cy)/fy, 1), row-major over pixels."""
        key = (self.width, self.height, self.focal_x, self.focal_y, self.center_x, self.center_y, through_pixel_center, str(device))
        if self._local_ray_directions_cache is not None and self._local_ray_directions_cache[0] == key:
            return self._local_ray_directions_cache[1]
        off = 0.5 if through_pixel_center else 0.0
        xs = torch.linspace((off - self.center_x) / self.focal_x, (self.width - 1 + off - self.center_x) / self.focal_x,
                            self.width, device=device)
        ys = torch.linspace((off - self.center_y) / self.focal_y, (self.height - 1 + off - self.center_y) / self.focal_y,
                            self.height, device=device)
        d = torch.empty((self.height, self.width, 3), dtype=torch.float32, device=device)
        d[..., 0] = xs[None, :]
        d[..., 1] = ys[:, None]
        d[..., 2] = 1.0
        d = d.reshape(-1, 3)
        self._local_ray_directions_cache = (key, d)
        return d
