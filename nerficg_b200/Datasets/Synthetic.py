"""Synthetic "Lego-shaped" dataset: Blender-synthetic geometry without the files.

Cameras sit on the upper hemisphere of radius 4.0311 looking at the origin with
camera_angle_x = 0.6911112070083618 (focal = 0.5/tan(fov/2) * W: 138.889 px at 100^2, 1111.11 px at 800^2),
principal point at the image centre, near 2, far 6, white background -- the geometry of
``dataset/nerf_synthetic/lego`` as read by the reference loader (src/Datasets/NeRF.py:45-107).  Targets are an
analytic scene (union of coloured axis-aligned boxes inside [-0.8, 0.8]^3) rendered exactly by ray/box
intersection: RGBA plus depth along the un-normalised ray, i.e. the quantities the NeRF path regresses.
The interface mirrors the reference BaseDataset (src/Datasets/Base.py): ``train()/test()/eval()`` switch the
active subset, ``ray_collection[mode]`` holds precomputed rays, ``default_camera`` carries the shared settings.
"""
from __future__ import annotations

import math

import torch

from .. import Framework
from ..Cameras.Perspective import PerspectiveCamera, SharedCameraSettings, fov_to_focal
from .utils import RayBatch, RayCollection, View

CAMERA_ANGLE_X = 0.6911112070083618
RADIUS = 4.031128874149275

# (centre, half extent, colour)
BOXES = [
    ((0.0, 0.0, -0.55), (0.75, 0.75, 0.08), (0.55, 0.55, 0.58)),
    ((-0.35, -0.30, -0.20), (0.25, 0.22, 0.28), (0.85, 0.15, 0.12)),
    ((0.35, 0.25, -0.10), (0.22, 0.30, 0.38), (0.95, 0.75, 0.10)),
    ((0.30, -0.40, -0.30), (0.18, 0.18, 0.18), (0.15, 0.35, 0.85)),
    ((-0.30, 0.40, -0.05), (0.15, 0.15, 0.42), (0.15, 0.65, 0.25)),
    ((0.00, 0.00, 0.35), (0.40, 0.10, 0.07), (0.80, 0.80, 0.85)),
    ((-0.55, -0.55, -0.35), (0.10, 0.10, 0.12), (0.60, 0.20, 0.70)),
]


def look_at_origin(theta: float, phi: float, radius: float = RADIUS) -> torch.Tensor:
    """4x4 camera-to-world: camera at spherical (theta azimuth, phi elevation), z forward to the origin, y down."""
    pos = radius * torch.tensor([math.cos(phi) * math.cos(theta), math.cos(phi) * math.sin(theta), math.sin(phi)])
    fwd = -pos / pos.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    c2w = torch.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, pos
    return c2w


@torch.no_grad()
def trace_scene(origin: torch.Tensor, direction: torch.Tensor):
    """Exact RGBA + depth of the box scene.  origin/direction: (n,3); depth is the ray parameter t (x = o + t d)."""
    n = origin.shape[0]
    best_t = torch.full((n,), float('inf'), device=origin.device)
    rgb = torch.zeros(n, 3, device=origin.device)
    inv = 1.0 / torch.where(direction.abs() < 1e-12, torch.full_like(direction, 1e-12), direction)
    light = torch.nn.functional.normalize(torch.tensor([0.4, 0.3, 0.85], device=origin.device), dim=0)
    for centre, half, colour in BOXES:
        c = torch.tensor(centre, device=origin.device)
        h = torch.tensor(half, device=origin.device)
        t0 = (c - h - origin) * inv
        t1 = (c + h - origin) * inv
        tmin, tmax = torch.minimum(t0, t1), torch.maximum(t0, t1)
        t_near, axis = tmin.max(dim=-1)
        t_far = tmax.min(dim=-1).values
        hit = (t_near <= t_far) & (t_near > 0) & (t_near < best_t)
        # Lambert shading by the face normal gives every box three distinguishable faces
        normal = torch.zeros(n, 3, device=origin.device)
        normal.scatter_(1, axis[:, None], -torch.sign(torch.gather(direction, 1, axis[:, None])))
        shade = (0.35 + 0.65 * (normal @ light).clamp(0, 1))[:, None]
        col = torch.tensor(colour, device=origin.device)[None, :] * shade
        rgb = torch.where(hit[:, None], col, rgb)
        best_t = torch.where(hit, t_near, best_t)
    alpha = torch.isfinite(best_t).float()[:, None]
    depth = torch.where(torch.isfinite(best_t), best_t, torch.zeros_like(best_t))[:, None]
    return rgb, alpha, depth


class SyntheticLegoDataset:
    """n_train / n_test / n_val posed views of the analytic scene at (width x height)."""

    def __init__(self, width: int = 800, height: int = 800, n_train: int = 100, n_test: int = 200, n_val: int = 0,
                 seed: int = 0, device: torch.device | str | None = None, background=(1.0, 1.0, 1.0),
                 near: float = 2.0, far: float = 6.0, render_targets: bool = True) -> None:
        self.device = torch.device(device) if device is not None else Framework.config.GLOBAL.get('DEFAULT_DEVICE', torch.device('cpu'))
        focal = fov_to_focal(CAMERA_ANGLE_X) * width
        self.shared = SharedCameraSettings(torch.tensor(background, dtype=torch.float32), near, far)
        self.default_camera = PerspectiveCamera(shared_settings=self.shared, width=width, height=height, focal_x=focal, focal_y=focal)
        g = torch.Generator().manual_seed(seed)
        self.data: dict[str, list[View]] = {}
        for mode, count in (('train', n_train), ('test', n_test), ('val', n_val)):
            views = []
            for i in range(count):
                if mode == 'test':  # smooth orbit like the Blender test trajectory
                    theta, phi = 2 * math.pi * i / max(count, 1), math.radians(30.0)
                else:
                    theta = float(torch.rand((), generator=g)) * 2 * math.pi
                    phi = math.asin(float(torch.rand((), generator=g)) * 0.95 + 0.02)
                views.append(View(self.default_camera, look_at_origin(theta, phi), frame_idx=i))
            self.data[mode] = views
        self.render_targets = render_targets
        self.ray_collection: dict[str, RayCollection | None] = {'train': None, 'test': None, 'val': None}
        self.mode = 'train'

    # ---- subset switching (reference Base.py) ----
    def train(self) -> 'SyntheticLegoDataset':
        self.mode = 'train'
        return self

    def test(self) -> 'SyntheticLegoDataset':
        self.mode = 'test'
        return self

    def eval(self) -> 'SyntheticLegoDataset':
        self.mode = 'val'
        return self

    def __len__(self) -> int:
        return len(self.data[self.mode])

    def __getitem__(self, index: int) -> View:
        view = self.data[self.mode][index]
        if self.render_targets and view.rgb is None:
            self._render_view(view)
        return view

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def _rays_of(self, view: View) -> tuple[torch.Tensor, torch.Tensor]:
        local = view.camera.compute_local_ray_directions(device=self.device)
        direction = local @ view.rotation.to(self.device).T
        return view.position.to(self.device).expand_as(direction).contiguous(), direction.contiguous()

    def _render_view(self, view: View) -> None:
        o, d = self._rays_of(view)
        rgb, alpha, depth = trace_scene(o, d)
        h, w = view.camera.height, view.camera.width
        view.rgb = rgb.reshape(h, w, 3).permute(2, 0, 1).contiguous()
        view.alpha = alpha.reshape(h, w, 1).permute(2, 0, 1).contiguous()
        view.depth = depth.reshape(h, w, 1).permute(2, 0, 1).contiguous()

    def precompute_rays(self, modes=('train',)) -> None:
        """All rays of the subsets as one RayCollection on the device (reference Base.py:172-216)."""
        keep = self.mode
        for mode in modes:
            batches, slices, start = [], [], 0
            for view in self.data[mode]:
                o, d = self._rays_of(view)
                rgb, alpha, depth = trace_scene(o, d) if self.render_targets else (torch.zeros_like(o), None, None)
                batches.append(RayBatch(origin=o, direction=d, view_direction=torch.nn.functional.normalize(d, dim=-1),
                                        rgb=rgb, alpha=alpha, depth=depth, _skip_post_init=True))
                slices.append(slice(start, start + o.shape[0]))
                start += o.shape[0]
            self.ray_collection[mode] = RayCollection(RayBatch.cat(batches), slices) if batches else None
        self.mode = keep

    def get_total_ray_count(self) -> int:
        return sum(v.camera.width * v.camera.height for v in self.data[self.mode])

    def get_all_rays(self) -> RayBatch:
        if self.ray_collection[self.mode] is None:
            self.precompute_rays([self.mode])
        return self.ray_collection[self.mode].all_rays
