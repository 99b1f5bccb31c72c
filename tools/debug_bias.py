import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nerficg_b200 import ops, params as P
from oracle import nerf_oracle as O
DEV = 'cuda:0'
sd = O.init_state_dict(0)
flat = P.flatten_state_dict(sd, 'nerf.', DEV)
packed = ops.mlp_pack(flat)
for n_rays, s in ((5, 77), (3, 128), (1, 130), (2, 64)):
    g = torch.Generator().manual_seed(7 * n_rays + s)
    o = torch.randn(n_rays, 3, generator=g) * 2
    d = torch.randn(n_rays, 3, generator=g)
    vd = torch.nn.functional.normalize(d, dim=-1)
    z = torch.sort(2 + 4 * torch.rand(n_rays, s, generator=g), -1).values
    n = n_rays * s
    up = (torch.randn(n, 4, generator=g) * 1e-3).to(DEV)
    scale = 2048.0
    stash = torch.empty(ops.mlp_stash_bytes(n), dtype=torch.uint8, device=DEV)
    out = ops.mlp_forward(packed, flat, o.to(DEV), d.to(DEV), vd.to(DEV), z.to(DEV), None, stash)
    upd = up.clone() * scale
    ws = torch.empty(ops.mlp_backward_workspace_bytes(n), dtype=torch.uint8, device=DEV)
    grads = torch.zeros_like(flat)
    ops.mlp_backward(grads, upd, out, stash, ws, packed, flat, n_rays, s, scale)
    torch.cuda.synchronize()
    v = P.views(grads)
    exp_bs = up[:, 3].half().float().sum().item()
    o4 = out.reshape(-1, 4)
    dpre = up[:, :3] * o4[:, :3] * (1 - o4[:, :3])
    print(n_rays, s, 'density bias got', v['density_layer.bias'].item(), 'expected', exp_bs,
          '| color bias got', v['color_layers.2.bias'].tolist(), 'expected', dpre.sum(0).tolist())
    per_tile = [up[i:i + 128, 3].sum().item() for i in range(0, n, 128)]
    print('   per-tile expected', per_tile)
