"""Spread of the CUDA-path test PSNR over repeated trainings (atomics make the weight gradients run-to-run
non-deterministic), next to the fp32 CPU oracle; plus a bitwise run-to-run check of the MLP forward / dgrad.
Development aid for tests/test_renderer_gpu.py::test_training_psnr_parity."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from oracle import nerf_oracle as O  # noqa: E402
from nerficg_b200 import Framework, ops, params  # noqa: E402

DEV = 'cuda:0'
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 120
n_rays, nc, nf = 256, 32, 64
Framework.setup(None, {'RENDERER.N_SAMPLES': nc + nf, 'RENDERER.COARSE_RATIO': nc / (nc + nf), 'RENDERER.RAY_BATCH_SIZE': 8192,
                       'TRAINING.NUM_ITERATIONS': 500000, 'GLOBAL.LOG_LEVEL': 0})
from nerficg_b200.Datasets.Synthetic import SyntheticLegoDataset  # noqa: E402
from nerficg_b200.Implementations import Methods  # noqa: E402
from nerficg_b200.Methods.NeRF import TRAINING_INSTANCE  # noqa: E402

# ---- bitwise determinism of the tensor-core chain ----
g = torch.Generator().manual_seed(0)
flat = (torch.rand(params.layout()[2], generator=g) - 0.5).mul(0.12).to(DEV)
packed = ops.mlp_pack(flat)
o = torch.randn(1000, 3, generator=g).to(DEV) * 0.1
d = torch.nn.functional.normalize(torch.randn(1000, 3, generator=g), dim=-1).to(DEV)
z = torch.sort(2 + 4 * torch.rand(1000, 77, generator=g), -1).values.to(DEV)
n = 1000 * 77
outs, gss = [], []
for _ in range(3):
    stash = torch.zeros(ops.mlp_stash_bytes(n), dtype=torch.uint8, device=DEV)
    ws = torch.zeros(ops.mlp_backward_workspace_bytes(n), dtype=torch.uint8, device=DEV)
    out = ops.mlp_forward(packed, flat, o, d, d, z, None, stash)
    up = torch.randn(n, 4, generator=torch.Generator(device=DEV).manual_seed(1), device=DEV) * 1e-3 * 1024
    ops.mlp_backward_dgrad(up, out, stash, ws, packed, flat, 1000, 77)
    torch.cuda.synchronize()
    outs.append((out.clone(), stash.clone()))
    gss.append(ws.clone())
print('forward bitwise repeatable:', all(torch.equal(outs[0][0], x[0]) and torch.equal(outs[0][1], x[1]) for x in outs[1:]))
print('dgrad   bitwise repeatable:', all(torch.equal(gss[0], x) for x in gss[1:]))

ds = SyntheticLegoDataset(48, 48, 6, 1, device='cpu')
ds.precompute_rays(['train', 'test'])
pool, test = ds.ray_collection['train'].all_rays, ds.ray_collection['test'].all_rays
bg = ds.default_camera.background_color
g = torch.Generator().manual_seed(0)
ids = [torch.randint(0, len(pool), (n_rays,), generator=g) for _ in range(steps)]
draws = [{'u_c': torch.rand(n_rays, nc, generator=g), 'u_f': torch.rand(n_rays, nf, generator=g)} for _ in range(steps)]
sd0 = O.init_state_dict(2)
gt = torch.lerp(bg.expand_as(test.rgb), test.rgb, test.alpha).clamp(0, 1)

sd = {k: v.clone().requires_grad_('frequency' not in k) for k, v in sd0.items()}
opt = torch.optim.Adam([v for k, v in sd.items() if v.requires_grad], lr=1.0)
for it in range(steps):
    b = pool[ids[it]]
    for grp in opt.param_groups:
        grp['lr'] = O.lr_factor(it, 5e-4, 5e-5, 500000)
    out = O.render_rays(sd, b.origin, b.direction, b.view_direction, 2.0, 6.0, bg, nc, nf, draws[it]['u_c'], draws[it]['u_f'])
    loss = O.nerf_loss(out, b.rgb, b.alpha, bg)
    opt.zero_grad()
    loss.backward()
    opt.step()
with torch.no_grad():
    ref = O.render_rays(sd, test.origin, test.direction, test.view_direction, 2.0, 6.0, bg, nc, nf)
print('oracle psnr', O.psnr(ref['rgb'].clamp(0, 1), gt))

for r in range(runs):
    model = Methods.get_model('NeRF', name='t')
    model.load_state_dict(sd0, strict=True)
    renderer = Methods.get_renderer('NeRF', model)
    trainer = TRAINING_INSTANCE(model=model, renderer=renderer)
    cam = ds.default_camera
    for it in range(steps):
        b = pool[ids[it]].to(device=torch.device(DEV))
        noise = [{k: v.to(DEV) for k, v in draws[it].items()}]
        out = renderer.render_rays(b, cam, randomize_samples=True, noise=noise)
        trainer.loss(out, b, bg.to(DEV)).backward()
        trainer.optimizer.step()
        trainer.optimizer.zero_grad()
        trainer.lr_scheduler.step()
    with torch.no_grad():
        got = renderer.render_rays(test.to(device=torch.device(DEV)), cam)
    print('cuda psnr run', r, O.psnr(got['rgb'].cpu().clamp(0, 1), gt))
