"""Fused backward (mlp_bwd_pipe.cu) against the legacy two-kernel path on identical inputs: per-tensor relative L2 and timings.
Development aid (run under `timeout`: a broken pipeline hand-shake would spin)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nerficg_b200 import ops, params  # noqa: E402

DEV = 'cuda:0'


def run(n_rays, s, time_it=False):
    g = torch.Generator().manual_seed(n_rays + s)
    flat = (torch.rand(params.layout()[2], generator=g) - 0.5).mul(0.12).to(DEV)
    packed = ops.mlp_pack(flat)
    o = torch.randn(n_rays, 3, generator=g).to(DEV) * 0.3
    d = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1).to(DEV)
    z = torch.sort(2 + 4 * torch.rand(n_rays, s, generator=g), -1).values.to(DEV)
    n = n_rays * s
    stash = torch.empty(ops.mlp_stash_bytes(n), dtype=torch.uint8, device=DEV)
    ws = torch.zeros(ops.mlp_backward_workspace_bytes(n), dtype=torch.uint8, device=DEV)
    out = ops.mlp_forward(packed, flat, o, d, d, z, None, stash)
    up = (torch.randn(n, 4, generator=g) * 1e-3 * 1024).to(DEV)
    up[:, 3] *= (out.reshape(-1, 4)[:, 3] > 0)
    g_old, g_new = torch.zeros_like(flat), torch.zeros_like(flat)
    ops.mlp_backward_legacy(g_old, up, out, stash, ws, packed, flat, n_rays, s, 1024.0)
    torch.cuda.synchronize()
    ops.mlp_backward_pipe(g_new, up, out, stash, ws, packed, flat, n_rays, s, 1024.0)
    torch.cuda.synchronize()
    worst = 0.0
    for (name, a), b in zip(params.views(g_new).items(), params.views(g_old).values()):
        rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
        worst = max(worst, rel)
        if rel > 2e-3 or not torch.isfinite(a).all():
            print(f'   {name:32s} rel {rel:.3e}  |new| {a.norm().item():.4e} |old| {b.norm().item():.4e}')
    print(f'{n_rays} x {s}: worst per-tensor rel L2 (pipe vs legacy) = {worst:.3e}', flush=True)
    if time_it:
        for name, fn in (('legacy', ops.mlp_backward_legacy), ('pipe', ops.mlp_backward_pipe)):
            for _ in range(2):
                fn(g_new, up, out, stash, ws, packed, flat, n_rays, s, 1024.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn(g_new, up, out, stash, ws, packed, flat, n_rays, s, 1024.0)
            b.record()
            torch.cuda.synchronize()
            print(f'   {name}: {a.elapsed_time(b) / 5:.3f} ms', flush=True)
    return worst


if __name__ == '__main__':
    sizes = [(2, 128), (5, 77), (64, 64), (300, 192), (2048, 64)]
    for n_rays, s in sizes:
        run(n_rays, s)
    run(4096, 64, True)
    run(4096, 192, True)
