"""Generates tests/golden/*.pt by running the UNMODIFIED reference on CPU.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference exists):
    python oracle/make_golden.py
The reference has no golden vectors of its own; these files pin
``oracle/nerf_oracle.py`` (and, on the GPU box, the CUDA path) to what nerficg's
own code computes on identical seeded inputs.  Weights are not stored: they come from
``nerf_oracle.init_state_dict(seed)`` (torch CPU generator) and a checksum is
stored instead.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import nerf_oracle as O  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

OUT = ROOT / 'tests' / 'golden'


def lego_rays(n_rays: int, seed: int, width: int = 100):
    """Random pixel rays of one Lego-shaped camera (radius 4.0311, fov 0.6911 rad)."""
    g = torch.Generator().manual_seed(seed)
    import math
    focal = 0.5 / math.tan(0.5 * 0.6911112070083618) * width
    th, ph = 0.9, 0.6
    pos = 4.0311 * torch.tensor([math.cos(ph) * math.cos(th), math.cos(ph) * math.sin(th), math.sin(ph)])
    fwd = -pos / pos.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    c2w = torch.stack((right, down, fwd, pos), dim=1)
    o, d, v = O.pinhole_rays(c2w, width, width, focal)
    ids = torch.randint(0, width * width, (n_rays,), generator=g)
    return o[ids].contiguous(), d[ids].contiguous(), v[ids].contiguous()


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    ref = load_reference(n_samples=192, coarse_ratio=0.3333333, seed=0)
    U, DS = ref['utils'], ref['ds_utils']
    method = ref['method']
    torch.manual_seed(1234)

    # ---- stage: stratified sampling (generate_samples) -------------------------------
    n, nc, nf = 12, 64, 128
    o, d, v = lego_rays(n, 1)
    rays = DS.RayBatch(origin=o, direction=d, view_direction=v)
    torch.manual_seed(11)
    u_c = torch.rand((n, nc))
    torch.manual_seed(11)
    z_rand = U.generate_samples(rays, nc, 2.0, 6.0, True)
    z_det = U.generate_samples(rays, nc, 2.0, 6.0, False)
    torch.save({'n': n, 'nc': nc, 'near': 2.0, 'far': 6.0, 'u': u_c, 'z_rand': z_rand, 'z_det': z_det.contiguous()},
               OUT / 'stratified.pt')

    # ---- stage: compositing (integrate_samples) + autograd ---------------------------
    s = 192
    g = torch.Generator().manual_seed(5)
    z = torch.sort(2.0 + 4.0 * torch.rand(n, s, generator=g), dim=-1).values
    z[:, 7] = z[:, 6]  # a tie as produced by the merge
    sigma = (torch.rand(n, s, generator=g) < 0.4).float() * (-torch.log(torch.rand(n, s, generator=g))) * 8.0
    sigma[0] = 0.0          # empty ray
    sigma[1, -1] = 3.0      # opaque last sample
    color = torch.rand(n, s, 3, generator=g)
    bg = torch.tensor([1.0, 1.0, 1.0])
    sig_r, col_r = sigma.clone().requires_grad_(True), color.clone().requires_grad_(True)
    rgb, depth, alpha, w = U.integrate_samples(z, d, sig_r, col_r, bg)
    g_rgb, g_alpha = torch.rand(n, 3, generator=g) - 0.5, torch.rand(n, 1, generator=g) - 0.5
    ((rgb * g_rgb).sum() + (alpha * g_alpha).sum()).backward()
    torch.save({'z': z, 'dirs': d, 'sigma': sigma, 'color': color, 'bg': bg, 'rgb': rgb.detach(), 'depth': depth.detach(),
                'alpha': alpha.detach(), 'w': w.detach(), 'g_rgb': g_rgb, 'g_alpha': g_alpha,
                'd_sigma': sig_r.grad, 'd_color': col_r.grad}, OUT / 'composite.pt')

    # ---- stage: importance sampling (generate_samples_from_pdf) + merge --------------
    w_c = w.detach()[:, :nc].contiguous()
    w_c[2] = 0.0  # zero weights -> uniform
    torch.manual_seed(13)
    u_f = torch.rand(n, nf)
    torch.manual_seed(13)
    zf_rand = U.generate_samples_from_pdf(z_rand, w_c, nf, True)
    zf_det = U.generate_samples_from_pdf(z_rand, w_c, nf, False)
    merged = torch.sort(torch.cat((z_rand, zf_rand), dim=-1), dim=-1)[0]
    torch.save({'z_coarse': z_rand, 'w_coarse': w_c, 'nf': nf, 'u': u_f, 'zf_rand': zf_rand, 'zf_det': zf_det,
                'merged_rand': merged}, OUT / 'importance.pt')

    # ---- stage: encoding + MLP (FrequencyEncoding, NeRFBlock) -----------------------
    sd = O.init_state_dict(seed=0)
    model = method.MODEL('golden').build()
    missing, unexpected = model.load_state_dict(sd, strict=True)
    npts = 96
    g = torch.Generator().manual_seed(7)
    pts = (torch.rand(npts, 3, generator=g) - 0.5) * 8.0
    vds = torch.nn.functional.normalize(torch.rand(npts, 3, generator=g) - 0.5, dim=-1)
    enc10 = model.nerf.encoding_position(pts)
    enc4 = model.nerf.encoding_direction(vds)
    sig_m, rgb_m = model.nerf(pts, vds)
    sig_cm, rgb_cm = model.coarse_nerf(pts, vds)
    kat = model.nerf.encoding_position(torch.tensor([[0.1, 0.2, 0.3]]))
    checksum = torch.stack([sd[k].double().abs().sum() for k in sorted(sd)]).sum()
    torch.save({'seed': 0, 'checksum': checksum, 'pts': pts, 'dirs': vds, 'enc_pos': enc10.detach(), 'enc_dir': enc4.detach(),
                'sigma': sig_m.detach(), 'rgb': rgb_m.detach(), 'sigma_coarse': sig_cm.detach(), 'rgb_coarse': rgb_cm.detach(),
                'kat_010203': kat.detach()}, OUT / 'mlp.pt')

    # ---- composed: render_rays + NeRFLoss + backward ---------------------------------
    nr = 24
    o, d, v = lego_rays(nr, 3)
    g = torch.Generator().manual_seed(9)
    rgb_gt, alpha_gt = torch.rand(nr, 3, generator=g), (torch.rand(nr, 1, generator=g) > 0.5).float()
    rays = DS.RayBatch(origin=o, direction=d, view_direction=v, rgb=rgb_gt, alpha=alpha_gt)
    cam = ref['PerspectiveCamera'](shared_settings=ref['SharedCameraSettings'](bg, 2.0, 6.0), width=100, height=100,
                                   focal_x=138.889, focal_y=138.889)
    renderer = method.RENDERER(model)
    assert (renderer.n_samples_coarse_nerf, renderer.n_samples_nerf) == (64, 128)
    renderer.RAY_BATCH_SIZE = 16  # two chunks (16 + 8) to pin the draw order across chunks
    noise_std = 0.5
    torch.manual_seed(21)
    draws = []
    for lo in range(0, nr, 16):
        m = min(16, nr - lo)
        draws.append({'u_c': torch.rand((m, 64)), 'n_c': noise_std * torch.randn(m * 64, 1),
                      'u_f': torch.rand(m, 128), 'n_f': noise_std * torch.randn(m * 192, 1)})
    torch.manual_seed(21)
    out = renderer.render_rays(rays, cam, randomize_samples=True, random_noise_density=noise_std)
    from Methods.NeRF.Loss import NeRFLoss
    loss_fn = NeRFLoss(1.0, 0.0, True)
    loss = loss_fn(out, rays, bg)
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    out_det = renderer.render_rays(rays, cam)  # deterministic, no noise
    torch.save({'seed': 0, 'o': o, 'd': d, 'v': v, 'rgb_gt': rgb_gt, 'alpha_gt': alpha_gt, 'bg': bg, 'chunk': 16,
                'noise_std': noise_std, 'draws': draws,
                'out': {k: t.detach() for k, t in out.items()}, 'loss': loss.detach(),
                'grad_norm': {k: t.norm() for k, t in grads.items()},
                'grad_head': {k: t.flatten()[:48].clone() for k, t in grads.items()},
                'out_det': {k: t.detach() for k, t in out_det.items()}}, OUT / 'render.pt')

    # ---- LR schedule (LRDecayPolicy) -------------------------------------------------
    from Optim.lr_utils import LRDecayPolicy
    pol = LRDecayPolicy(lr_init=5e-4, lr_final=5e-5, max_steps=500000)
    its = [0, 1, 1000, 250000, 499999, 500000, 600000]
    torch.save({'its': its, 'lr': [pol(i) for i in its]}, OUT / 'lr.pt')
    for f in sorted(OUT.glob('*.pt')):
        print(f.name, f.stat().st_size)


if __name__ == '__main__':
    main()
