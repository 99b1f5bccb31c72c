"""Micro-timings of the C-ABI kernels with CUDA events (development aid; bench.py is the contract)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from nerficg_b200 import ops, params  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402

DEV = 'cuda:0'


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


def main():
    sd = O.init_state_dict(0)
    flat = params.flatten_state_dict(sd, 'nerf.', DEV)
    packed = ops.mlp_pack(flat)
    for n_rays, s in ((4096, 64), (4096, 192), (65536, 192)):
        g = torch.Generator().manual_seed(0)
        o = torch.randn(n_rays, 3, generator=g).to(DEV)
        d = torch.randn(n_rays, 3, generator=g).to(DEV)
        vd = torch.nn.functional.normalize(d, dim=-1)
        z = torch.sort(2 + 4 * torch.rand(n_rays, s, generator=g), -1).values.to(DEV)
        n = n_rays * s
        fl_f, fl_b = 1186816 * n, 2302208 * n
        t = timeit(lambda: ops.mlp_forward(packed, flat, o, d, vd, z))
        print(f'fwd infer  {n_rays}x{s}: {t:.3f} ms  {fl_f / t / 1e9:.1f} TFLOP/s')
        stash = torch.empty(ops.mlp_stash_bytes(n), dtype=torch.uint8, device=DEV)
        t = timeit(lambda: ops.mlp_forward(packed, flat, o, d, vd, z, None, stash))
        print(f'fwd train  {n_rays}x{s}: {t:.3f} ms  {fl_f / t / 1e9:.1f} TFLOP/s')
        out = ops.mlp_forward(packed, flat, o, d, vd, z, None, stash)
        ws = torch.empty(ops.mlp_backward_workspace_bytes(n), dtype=torch.uint8, device=DEV)
        up = (torch.randn(n, 4, device=DEV) * 1e-3 * 1024)
        grads = torch.zeros_like(flat)
        t = timeit(lambda: ops.mlp_backward(grads, up, out, stash, ws, packed, flat, n_rays, s, 1024.0))
        print(f'bwd        {n_rays}x{s}: {t:.3f} ms  {fl_b / t / 1e9:.1f} TFLOP/s')
        rs = out
        dirs = d
        bg = torch.ones(3, device=DEV)
        t = timeit(lambda: ops.composite_forward(z, rs, dirs, bg, True))
        print(f'composite fwd {n_rays}x{s}: {t * 1e3:.1f} us  {(n * 24 + n_rays * 32) / t / 1e6:.1f} GB/s')
        grgb = torch.randn(n_rays, 3, device=DEV)
        t = timeit(lambda: ops.composite_backward(z, rs, dirs, bg, grgb, None, None, True, 1024.0))
        print(f'composite bwd {n_rays}x{s}: {t * 1e3:.1f} us  {(n * 36 + n_rays * 48) / t / 1e6:.1f} GB/s')
    t = timeit(lambda: ops.mlp_pack(flat, packed))
    print(f'pack: {t * 1e3:.1f} us')


if __name__ == '__main__':
    main()
