"""CPU oracle for the vanilla-NeRF volume-rendering hot path.  TEST INFRASTRUCTURE ONLY.

This file is an independent fp32 PyTorch restatement of the algorithm of
nerficg's ``src/Methods/NeRF`` (reference checkout: /root/reference).  It is the
checker for the CUDA path, never the thing that is shipped or measured: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The product package
``nerficg_b200`` must not import anything from ``oracle/``.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so this restatement is pinned against outputs of the
reference code itself, generated in the build container by
``oracle/make_golden.py`` (which imports /root/reference through
``oracle/ref_loader.py``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` replays them.

Every function cites the reference lines it restates (paths relative to
/root/reference).  All tensors are fp32 unless noted.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

FINAL_DELTA = 1.0e10  # src/Methods/NeRF/utils.py:118
PDF_EPS = 1.0e-5      # src/Methods/NeRF/utils.py:86,106


# --------------------------------------------------------------------------------------
# sampling
# --------------------------------------------------------------------------------------
def stratified_depths(n_rays: int, n_samples: int, near: float, far: float,
                      u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Depth samples per ray, (n_rays, n_samples).

    Restates ``generate_samples`` (src/Methods/NeRF/utils.py:57-75): an fp32
    ``linspace(near, far, n)`` shared by all rays; when uniform noise ``u`` of
    shape (n_rays, n_samples) is supplied, every linspace node becomes a stratum
    whose borders are the midpoints to its neighbours (the first and last strata
    are half width) and the sample is ``lo + (hi - lo) * u``.
    """
    t = torch.linspace(near, far, n_samples, dtype=torch.float32)
    if u is None:
        return t.expand(n_rays, n_samples).clone()
    mid = 0.5 * (t[1:] + t[:-1])
    hi = torch.cat((mid, t[-1:]))
    lo = torch.cat((t[:1], mid))
    return lo[None, :] + (hi - lo)[None, :] * u.to(torch.float32)


def importance_depths(z_coarse: torch.Tensor, weights: torch.Tensor, n_fine: int,
                      u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Inverse-CDF samples, (n_rays, n_fine); no gradient.

    Restates ``generate_samples_from_pdf`` (src/Methods/NeRF/utils.py:78-109):
    bin edges are midpoints of the coarse depths (Nc-1 edges); the pdf is built
    from the interior weights w[1:-1] + 1e-5 (Nc-2 values); cdf = [0, cumsum];
    ``u`` is uniform noise (n_rays, n_fine) or, when None, linspace(0, 1, n_fine);
    index = number of cdf entries <= u (searchsorted right=True); a bracket whose
    cdf span is < 1e-5 uses denominator 1.
    """
    edges = 0.5 * (z_coarse[:, :-1] + z_coarse[:, 1:])
    v = weights[:, 1:-1] + PDF_EPS
    pdf = v / v.sum(dim=-1, keepdim=True)
    cdf = torch.cat((torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, dim=-1)), dim=-1)
    if u is None:
        u = torch.linspace(0.0, 1.0, n_fine, dtype=torch.float32).expand(cdf.shape[0], n_fine)
    u = u.contiguous()
    idx = torch.searchsorted(cdf, u, right=True)
    lo = (idx - 1).clamp(min=0)
    hi = idx.clamp(max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    b_lo, b_hi = torch.gather(edges, 1, lo), torch.gather(edges, 1, hi)
    span = c_hi - c_lo
    span = torch.where(span < PDF_EPS, torch.ones_like(span), span)
    frac = (u - c_lo) / span
    return (b_lo + frac * (b_hi - b_lo)).detach()


def merge_depths(z_coarse: torch.Tensor, z_fine: torch.Tensor) -> torch.Tensor:
    """Ascending per-ray union of both sample sets (src/Methods/NeRF/Renderer.py:70)."""
    return torch.sort(torch.cat((z_coarse, z_fine), dim=-1), dim=-1).values


# --------------------------------------------------------------------------------------
# encoding + MLP
# --------------------------------------------------------------------------------------
def frequency_encoding(x: torch.Tensor, n_freq: int, append_input: bool = True) -> torch.Tensor:
    """[x, per coordinate: cos(2^k x_i) for k<L, then sin(2^k x_i) for k<L].

    Restates ``FrequencyEncoding`` (src/Methods/NeRF/utils.py:12-36): factors are
    2^k with no pi; the cos block precedes the sin block; coordinate-major.
    """
    f = torch.exp2(torch.arange(n_freq, dtype=torch.float32))
    arg = x[..., None] * f
    enc = torch.cat((torch.cos(arg), torch.sin(arg)), dim=-1).flatten(start_dim=-2)
    return torch.cat((x, enc), dim=-1) if append_input else enc


def mlp_forward(sd: dict, prefix: str, positions: torch.Tensor, directions: torch.Tensor,
                noise: Optional[torch.Tensor] = None, n_freq_pos: int = 10, n_freq_dir: int = 4,
                n_layers: int = 8, skips=(5,), return_intermediates: bool = False, operand_dtype=None):
    """Density (N,1) and colour (N,3) of one NeRF block.

    Restates ``NeRFBlock.forward`` (src/Methods/NeRF/Model.py:59-83) functionally
    on a state dict (keys as dumped in SURVEY.md A.6, ``prefix`` = 'nerf.' or
    'coarse_nerf.').  The encoded position is concatenated *after* the hidden
    features before the layers listed in ``skips``; density = ReLU(linear + noise);
    colour = sigmoid(W2 ReLU(W1 [feature, enc(dir)])).  ``noise`` is the already
    scaled additive density noise (std * randn) or None.

    ``operand_dtype`` (e.g. torch.float16) is NOT part of the reference: it rounds the operands of
    the ten tensor-core GEMMs (hidden layers, feature layer, first colour layer) the way the CUDA
    path does (fp32 accumulate, fp32 heads), so tests can separate operand rounding from real bugs.
    """
    def lin(x, w, b):
        if operand_dtype is not None:  # straight-through rounding: gradients stay fp32
            x = x + (x.to(operand_dtype).float() - x).detach()
            w = w + (w.to(operand_dtype).float() - w).detach()
        return torch.nn.functional.linear(x, w, b)

    ex = frequency_encoding(positions, n_freq_pos)
    h = ex
    inter = {}
    for l in range(n_layers):
        w, b = sd[f'{prefix}initial_layers.{l}.0.weight'], sd[f'{prefix}initial_layers.{l}.0.bias']
        h = torch.relu(lin(h, w, b))
        inter[f'h{l}'] = h
        if (l + 1) in skips:
            h = torch.cat((h, ex), dim=-1)
    raw = torch.nn.functional.linear(h, sd[f'{prefix}density_layer.weight'], sd[f'{prefix}density_layer.bias'])
    if noise is not None:
        raw = raw + noise
    sigma = torch.relu(raw)
    ed = frequency_encoding(directions, n_freq_dir)
    feat = lin(h, sd[f'{prefix}feature_layer.weight'], sd[f'{prefix}feature_layer.bias'])
    g = torch.relu(lin(torch.cat((feat, ed), dim=-1), sd[f'{prefix}color_layers.0.weight'], sd[f'{prefix}color_layers.0.bias']))
    rgb = torch.sigmoid(torch.nn.functional.linear(g, sd[f'{prefix}color_layers.2.weight'], sd[f'{prefix}color_layers.2.bias']))
    if return_intermediates:
        inter.update(feat=feat, g=g, raw=raw)
        return sigma, rgb, inter
    return sigma, rgb


def init_state_dict(seed: int = 0, hierarchical: bool = True) -> dict:
    """Random parameters with torch's default ``nn.Linear`` init and the reference's
    key names/shapes (src/Methods/NeRF/Model.py:35-54)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(key, n_out, n_in):
        bound = 1.0 / math.sqrt(n_in)
        sd[key + '.weight'] = (torch.rand(n_out, n_in, generator=g) * 2 - 1) * bound
        sd[key + '.bias'] = (torch.rand(n_out, generator=g) * 2 - 1) * bound

    for prefix in (['coarse_nerf.'] if hierarchical else []) + ['nerf.']:
        sd[prefix + 'encoding_position.frequency_factors'] = torch.exp2(torch.arange(10.0))[None, None, :]
        sd[prefix + 'encoding_direction.frequency_factors'] = torch.exp2(torch.arange(4.0))[None, None, :]
        for l in range(8):
            lin(f'{prefix}initial_layers.{l}.0', 256, 63 if l == 0 else (319 if l == 5 else 256))
        lin(prefix + 'feature_layer', 256, 256)
        lin(prefix + 'density_layer', 1, 256)
        lin(prefix + 'color_layers.0', 128, 283)
        lin(prefix + 'color_layers.2', 3, 128)
    return sd


# --------------------------------------------------------------------------------------
# compositing
# --------------------------------------------------------------------------------------
def composite(z: torch.Tensor, ray_dirs: torch.Tensor, sigma: torch.Tensor, color: torch.Tensor,
              background: Optional[torch.Tensor]):
    """rgb (n,3), depth (n,1), alpha (n,1), weights (n,S).

    Restates ``integrate_samples`` (src/Methods/NeRF/utils.py:112-136): interval
    lengths are z differences with a last interval of 1e10, all scaled by the norm
    of the (un-normalised) ray direction; alpha_i = 1-exp(-sigma_i delta_i);
    transmittance by cumulative product; depth = sum(w z)/alpha where the final
    transmittance is < 1, else 0; colour gets final transmittance x background.
    """
    delta = torch.cat((z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], FINAL_DELTA)), dim=-1)
    delta = delta * ray_dirs.norm(dim=-1, keepdim=True)
    a = 1.0 - torch.exp(-sigma * delta)
    trans = torch.cumprod(torch.cat((torch.ones_like(a[:, :1]), 1.0 - a), dim=-1), dim=-1)
    w = a * trans[:, :-1]
    t_final = trans[:, -1:]
    alpha = 1.0 - t_final
    depth = torch.where(t_final < 1.0, (w * z).sum(dim=-1, keepdim=True) / alpha, torch.zeros_like(alpha))
    rgb = (w[..., None] * color).sum(dim=-2)
    if background is not None:
        rgb = rgb + t_final * background
    return rgb, depth, alpha, w


# --------------------------------------------------------------------------------------
# composed renderer, loss, schedule
# --------------------------------------------------------------------------------------
def render_rays(sd: dict, origins: torch.Tensor, dirs: torch.Tensor, view_dirs: torch.Tensor,
                near: float, far: float, background: torch.Tensor, n_coarse: int, n_fine: int,
                u_coarse: Optional[torch.Tensor] = None, u_fine: Optional[torch.Tensor] = None,
                noise_coarse: Optional[torch.Tensor] = None, noise_fine: Optional[torch.Tensor] = None,
                z_override: Optional[torch.Tensor] = None) -> dict:
    """One chunk of ``NeRFRayRenderingComponent.forward`` (src/Methods/NeRF/Renderer.py:29-95).

    n_coarse == 0 selects the single-pass branch (line 72).  ``z_override``
    teacher-forces the fine-pass sample positions (used for stage-wise parity).
    """
    n = origins.shape[0]
    out = {}
    if n_coarse > 0:
        zc = stratified_depths(n, n_coarse, near, far, u_coarse)
        xc = origins[:, None, :] + dirs[:, None, :] * zc[:, :, None]
        vd = view_dirs[:, None, :].expand_as(xc)
        s_c, c_c = mlp_forward(sd, 'coarse_nerf.', xc.reshape(-1, 3), vd.reshape(-1, 3), noise_coarse)
        rgb_c, depth_c, alpha_c, w_c = composite(zc, dirs, s_c.reshape(n, n_coarse), c_c.reshape(n, n_coarse, 3), background)
        zf = importance_depths(zc, w_c, n_fine, u_fine)
        z = merge_depths(zc, zf)
        out.update(rgb_coarse=rgb_c, depth_coarse=depth_c, alpha_coarse=alpha_c, z_coarse=zc, w_coarse=w_c, z_fine=zf)
    else:
        z = stratified_depths(n, n_fine, near, far, u_fine)
    if z_override is not None:
        z = z_override
    s_tot = n_coarse + n_fine
    x = origins[:, None, :] + dirs[:, None, :] * z[:, :, None]
    vd = view_dirs[:, None, :].expand_as(x)
    s_f, c_f = mlp_forward(sd, 'nerf.', x.reshape(-1, 3), vd.reshape(-1, 3), noise_fine)
    rgb, depth, alpha, w = composite(z, dirs, s_f.reshape(n, s_tot), c_f.reshape(n, s_tot, 3), background)
    out.update(rgb=rgb, depth=depth, alpha=alpha, z=z, w=w)
    return out


def nerf_loss(outputs: dict, rgb_gt: torch.Tensor, alpha_gt: Optional[torch.Tensor], background: torch.Tensor,
              lambda_color: float = 1.0, lambda_alpha: float = 0.0) -> torch.Tensor:
    """Restates ``NeRFLoss.forward`` (src/Methods/NeRF/Loss.py:26-43) with
    ``apply_background_color`` (src/Datasets/utils.py:185-189): the target is
    clamp(lerp(bg, rgb_gt, alpha_gt), 0, 1); MSE on fine and coarse colour (+ alpha terms)."""
    if alpha_gt is None:
        alpha_gt = torch.ones_like(outputs['alpha'])
    target = torch.lerp(background.expand_as(rgb_gt), rgb_gt, alpha_gt).clamp(0, 1)
    mse = torch.nn.functional.mse_loss
    loss = lambda_color * mse(outputs['rgb'], target)
    if lambda_alpha > 0:
        loss = loss + lambda_alpha * mse(outputs['alpha'], alpha_gt)
    if 'rgb_coarse' in outputs:
        loss = loss + lambda_color * mse(outputs['rgb_coarse'], target)
        if lambda_alpha > 0:
            loss = loss + lambda_alpha * mse(outputs['alpha_coarse'], alpha_gt)
    return loss


def lr_factor(iteration: int, lr_init: float, lr_final: float, max_steps: int) -> float:
    """Log-linear decay (src/Optim/lr_utils.py:18-32, delay disabled as in NeRF/Trainer.py:35)."""
    if iteration < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    t = min(max(iteration / max_steps, 0.0), 1.0)
    return float(math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t))


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    """-10 log10(mse) with data range 1 (torchmetrics PSNR as used in NeRF/Loss.py:19; unpinned dependency)."""
    return float(-10.0 * torch.log10(torch.mean((a - b) ** 2)))


# --------------------------------------------------------------------------------------
# synthetic "Lego-shaped" rays (mirrors nerficg_b200.synthetic but kept independent)
# --------------------------------------------------------------------------------------
def pinhole_rays(c2w: torch.Tensor, width: int, height: int, focal: float):
    """Pixel-centre rays: origin, un-normalised direction (camera z = 1), unit view direction.

    Restates ``PerspectiveCamera.compute_local_ray_directions``
    (src/Cameras/Perspective.py:64-94) + ``View.get_rays`` (src/Datasets/utils.py:1053-1074).
    ``c2w`` is a 3x4 matrix whose rotation columns map camera axes to world.
    """
    cx, cy = width / 2, height / 2
    xs = torch.linspace((0.5 - cx) / focal, (width - 0.5 - cx) / focal, width)
    ys = torch.linspace((0.5 - cy) / focal, (height - 0.5 - cy) / focal, height)
    local = torch.stack((xs[None, :].expand(height, width), ys[:, None].expand(height, width),
                         torch.ones(height, width)), dim=-1).reshape(-1, 3)
    d = local @ c2w[:3, :3].T
    o = c2w[:3, 3].expand_as(d)
    return o.contiguous(), d.contiguous(), torch.nn.functional.normalize(d, dim=-1)


def camera_rays(c2w: torch.Tensor, width: int, height: int, focal_x: float, focal_y: float, center_x: float, center_y: float,
                pixel_ids: Optional[torch.Tensor] = None):
    """General pinhole camera (fx != fy, off-centre principal point): restates ``compute_local_ray_directions``
    (src/Cameras/Perspective.py:64-94, distortion-free) + ``View.cam_to_world`` / ``View.get_rays``
    (src/Datasets/utils.py:1033-1038, 1053-1074).  ``c2w``: 3x4 / 4x4 (float64 as the reference stores it; rotation and
    position are cast to float32 like ``View.rotation`` / ``View.position``).  Returns origin, direction, view_direction."""
    xs = torch.linspace((0.5 - center_x) / focal_x, (width - 1 + 0.5 - center_x) / focal_x, width)
    ys = torch.linspace((0.5 - center_y) / focal_y, (height - 1 + 0.5 - center_y) / focal_y, height)
    local = torch.empty((height, width, 3), dtype=torch.float32)
    local[..., 0] = xs[None, :]
    local[..., 1] = ys[:, None]
    local[..., 2] = 1.0
    local = local.reshape(-1, 3)
    rot = c2w[:3, :3].to(torch.float32)
    d = local @ rot.T
    o = c2w[:3, 3].to(torch.float32).expand_as(d)
    v = torch.nn.functional.normalize(d, dim=-1)
    if pixel_ids is not None:
        o, d, v = o[pixel_ids], d[pixel_ids], v[pixel_ids]
    return o.contiguous(), d.contiguous(), v.contiguous()
