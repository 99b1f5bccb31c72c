"""Trained-regime golden vectors from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY; build container only.

    python oracle/make_golden_trained.py [--steps 900] [--batch 512]

Every other golden in tests/golden/ uses freshly initialised weights (sigma ~ 0, rgb ~ 0.5), where fp16 operand
rounding is harmless.  This script trains the reference's own NeRF (its Model / Renderer / NeRFLoss / Adam x LambdaLR,
src/Methods/NeRF/Trainer.py:32-63) on CPU on the analytic box scene until densities are sharp, then writes

  * tests/golden/ref_trained_checkpoint.pt  -- the checkpoint file written by the reference's ``BaseModel.save``
    (src/Methods/Base/Model.py:103-111): the reference-written checkpoint our ``NeRF.load`` must read;
  * tests/golden/trained_render.pt          -- for 256 held-out rays: the reference's deterministic ``render_rays``
    outputs, a randomised run with the noise it drew, the per-sample (sigma, rgb) of both blocks at the reference's own
    sample positions (teacher forcing for K3), and loss + parameter gradients of one training batch at these weights.
"""
from __future__ import annotations

import argparse
import math
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import nerf_oracle as O  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

OUT = ROOT / 'tests' / 'golden'
CAMERA_ANGLE_X = 0.6911112070083618
RADIUS = 4.031128874149275


def camera_rays(theta: float, phi: float, width: int):
    focal = 0.5 / math.tan(0.5 * CAMERA_ANGLE_X) * width
    pos = RADIUS * torch.tensor([math.cos(phi) * math.cos(theta), math.cos(phi) * math.sin(theta), math.sin(phi)])
    fwd = -pos / pos.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    c2w = torch.stack((right, down, fwd, pos), dim=1)
    return O.pinhole_rays(c2w, width, width, focal)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=900)
    ap.add_argument('--batch', type=int, default=512)
    ap.add_argument('--width', type=int, default=100)
    ap.add_argument('--lr-steps', type=int, default=500000, help='max_steps of the LR policy (the shipped 500k keeps lr ~ 5e-4 over a short run)')
    args = ap.parse_args()

    ref = load_reference(n_samples=192, coarse_ratio=0.3333333, seed=0)
    RF, DS, method, U = ref['Framework'], ref['ds_utils'], ref['method'], ref['utils']
    from nerficg_b200.Datasets.Synthetic import trace_scene          # analytic targets (data only)
    from Methods.NeRF.Loss import NeRFLoss
    from Optim.lr_utils import LRDecayPolicy

    bg = torch.tensor([1.0, 1.0, 1.0])
    cam = ref['PerspectiveCamera'](shared_settings=ref['SharedCameraSettings'](bg, 2.0, 6.0), width=args.width, height=args.width,
                                   focal_x=0.5 / math.tan(0.5 * CAMERA_ANGLE_X) * args.width, focal_y=0.5 / math.tan(0.5 * CAMERA_ANGLE_X) * args.width)
    g = torch.Generator().manual_seed(0)
    pool = []
    for _ in range(24):                                                # 24 training views on the upper hemisphere
        theta = float(torch.rand((), generator=g)) * 2 * math.pi
        phi = math.asin(float(torch.rand((), generator=g)) * 0.95 + 0.02)
        o, d, v = camera_rays(theta, phi, args.width)
        rgb, alpha, _ = trace_scene(o, d)
        pool.append((o, d, v, rgb, alpha))
    o_all, d_all, v_all, rgb_all, a_all = (torch.cat(t) for t in zip(*pool))

    torch.manual_seed(0)
    model = method.MODEL('ref_trained').build()
    renderer = method.RENDERER(model)
    assert (renderer.n_samples_coarse_nerf, renderer.n_samples_nerf) == (64, 128)
    optimizer = torch.optim.Adam(model.parameters(), lr=1.0)
    sched = torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=LRDecayPolicy(lr_init=5e-4, lr_final=5e-5, max_steps=args.lr_steps))
    loss_fn = NeRFLoss(1.0, 0.0, True)
    t0 = time.time()
    for it in range(args.steps):
        ids = torch.randint(0, o_all.shape[0], (args.batch,), generator=g)
        rays = DS.RayBatch(origin=o_all[ids], direction=d_all[ids], view_direction=v_all[ids], rgb=rgb_all[ids], alpha=a_all[ids])
        out = renderer.render_rays(rays, cam, randomize_samples=True, random_noise_density=0.0)
        loss = loss_fn(out, rays, bg)
        loss.backward()
        optimizer.step()
        optimizer.zero_grad()
        sched.step()
        model.num_iterations_trained += 1
        if it % 20 == 0 or it == args.steps - 1:
            print(f'it {it:4d} loss {float(loss):.5f}  psnr {-10 * math.log10(float(loss) / 2 + 1e-12):.2f} dB  {time.time() - t0:.0f}s', flush=True)

    model.eval()
    model.save(OUT / 'ref_trained_checkpoint.pt')                       # the reference's own writer
    sd = {k: t.detach().clone() for k, t in model.state_dict().items()}

    # ---- held-out rays: deterministic + randomised render, teacher-forced per-sample outputs ----
    n = 256
    o, d, v = camera_rays(0.35, math.radians(33.0), args.width)
    gt_rgb, gt_alpha, _ = trace_scene(o, d)
    ids = torch.randperm(o.shape[0], generator=g)[:n]
    o, d, v, gt_rgb, gt_alpha = o[ids].contiguous(), d[ids].contiguous(), v[ids].contiguous(), gt_rgb[ids].contiguous(), gt_alpha[ids].contiguous()
    rays = DS.RayBatch(origin=o, direction=d, view_direction=v, rgb=gt_rgb, alpha=gt_alpha)
    with torch.no_grad():
        out_det = renderer.render_rays(rays, cam)
        # stage-wise, deterministic: what render_rays computed inside (src/Methods/NeRF/Renderer.py:50-83)
        z_c = U.generate_samples(rays, 64, 2.0, 6.0, False).expand(n, 64).contiguous()
        pts_c = (o[:, None, :] + d[:, None, :] * z_c[:, :, None]).reshape(-1, 3)
        vd_c = v[:, None, :].expand(n, 64, 3).reshape(-1, 3)
        sig_c, rgb_c = model.coarse_nerf(pts_c, vd_c)
        _, _, _, w_c = U.integrate_samples(z_c, d, sig_c.reshape(n, 64), rgb_c.reshape(n, 64, 3), bg)
        z_f = U.generate_samples_from_pdf(z_c, w_c, 128, False)
        z = torch.sort(torch.cat((z_c, z_f), dim=-1), dim=-1)[0]
        pts = (o[:, None, :] + d[:, None, :] * z[:, :, None]).reshape(-1, 3)
        vd = v[:, None, :].expand(n, 192, 3).reshape(-1, 3)
        sig_f, rgb_f = model.nerf(pts, vd)
        rgb_tf, depth_tf, alpha_tf, _ = U.integrate_samples(z, d, sig_f.reshape(n, 192), rgb_f.reshape(n, 192, 3), bg)
        assert torch.allclose(rgb_tf, out_det['rgb'], atol=1e-6), 'stage replay differs from render_rays'
    torch.manual_seed(31)
    draws = [{'u_c': torch.rand((n, 64)), 'u_f': torch.rand(n, 128)}]
    torch.manual_seed(31)
    renderer.RAY_BATCH_SIZE = 8192
    model.train()
    out_rand = renderer.render_rays(rays, cam, randomize_samples=True, random_noise_density=0.0)
    loss = loss_fn(out_rand, rays, bg)
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    psnr = float(-10 * torch.log10(torch.mean((out_det['rgb'] - DS.apply_background_color(gt_rgb, gt_alpha, bg, is_chw=False)) ** 2)))
    torch.save({
        'steps': args.steps, 'batch': args.batch, 'psnr_heldout': psnr, 'sigma_max': float(sig_f.max()),
        'o': o, 'd': d, 'v': v, 'rgb_gt': gt_rgb, 'alpha_gt': gt_alpha, 'bg': bg,
        'out_det': {k: t.detach() for k, t in out_det.items()},
        'z_c': z_c, 'sigma_c': sig_c.reshape(n, 64), 'rgb_c': rgb_c.reshape(n, 64, 3), 'w_c': w_c,
        'z': z, 'sigma_f': sig_f.reshape(n, 192), 'rgb_f': rgb_f.reshape(n, 192, 3),
        'draws': draws, 'out_rand': {k: t.detach() for k, t in out_rand.items()}, 'loss': loss.detach(),
        'grad_norm': {k: t.norm() for k, t in grads.items()},
        'grad_head': {k: t.flatten()[:256].clone() for k, t in grads.items()},
        'checksum': torch.stack([sd[k].double().abs().sum() for k in sorted(sd)]).sum(),
    }, OUT / 'trained_render.pt')
    print(f'held-out PSNR {psnr:.2f} dB, sigma_max {float(sig_f.max()):.1f}')
    for f in ('ref_trained_checkpoint.pt', 'trained_render.pt'):
        print(f, (OUT / f).stat().st_size)


if __name__ == '__main__':
    main()
