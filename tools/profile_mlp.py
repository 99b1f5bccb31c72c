"""Launches each MLP kernel a few times at the bench size (for ncu -k regex:...)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nerficg_b200 import ops, params  # noqa: E402
DEV = 'cuda:0'
n_rays, s = 4096, 192
g = torch.Generator().manual_seed(0)
flat = (torch.rand(params.layout()[2], generator=g) - 0.5).mul(0.12).to(DEV)
packed = ops.mlp_pack(flat)
o = torch.randn(n_rays, 3, generator=g).to(DEV) * 0.1
d = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1).to(DEV)
z = torch.sort(2 + 4 * torch.rand(n_rays, s, generator=g), -1).values.to(DEV)
n = n_rays * s
stash = torch.empty(ops.mlp_stash_bytes(n), dtype=torch.uint8, device=DEV)
ws = torch.empty(ops.mlp_backward_workspace_bytes(n), dtype=torch.uint8, device=DEV)
grads = torch.zeros_like(flat)
up = torch.randn(n, 4, device=DEV) * 1e-3 * 1024
for _ in range(3):
    ops.mlp_forward(packed, flat, o, d, d, z)                       # mlp_fwd_kernel<false>
    out = ops.mlp_forward(packed, flat, o, d, d, z, None, stash)    # mlp_fwd_kernel<true>
    ops.mlp_backward_dgrad(up, out, stash, ws, packed, flat, n_rays, s)
    ops.mlp_backward_wgrad(grads, stash, ws, n_rays, s, 1024.0)
torch.cuda.synchronize()
