"""Stall accounting of the fused backward pipeline (mlp_bwd_pipe.cu): where each role of a stage pair spends its cycles.
Averages per stage-leader CTA (64 of them) / per head CTA slot thread (32), in cycles."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nerficg_b200 import _lib, ops, params  # noqa: E402
import os
if os.environ.get('NERF_B200_LIB'):
    _lib.LIB_PATH = Path(os.environ['NERF_B200_LIB']).resolve()

DEV = 'cuda:0'
NAMES = {0: 'loader wait ready flag', 1: 'loader wait ring empty', 2: 'mma wait acc free', 3: 'mma wait operands (dgrad)', 4: 'mma wait operands (wgrad)', 5: 'mma total',
         6: 'epilogue wait acc ready', 7: 'epilogue wait image copied', 8: 'epilogue total', 9: 'store wait image', 10: 'store wait slot freed',
         11: 'store copy + fence', 12: 'store total', 13: 'head epilogue total', 14: 'head wait slot freed', 15: 'head wait store done',
         16: 'head wait acc ready', 21: 'loader proxy fence', 22: 'loader total', 23: 'loader issue', 24: 'store fence + flag'}


def main():
    n_rays, s = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 192)
    g = torch.Generator().manual_seed(0)
    flat = (torch.rand(params.layout()[2], generator=g) - 0.5).mul(0.12).to(DEV)
    packed = ops.mlp_pack(flat)
    o = torch.randn(n_rays, 3, generator=g).to(DEV) * 0.1
    d = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1).to(DEV)
    z = torch.sort(2 + 4 * torch.rand(n_rays, s, generator=g), -1).values.to(DEV)
    n = n_rays * s
    stash = torch.empty(ops.mlp_stash_bytes(n), dtype=torch.uint8, device=DEV)
    ws = torch.zeros(ops.mlp_backward_workspace_bytes(n), dtype=torch.uint8, device=DEV)
    out = ops.mlp_forward(packed, flat, o, d, d, z, None, stash)
    up = torch.randn(n, 4, device=DEV) * 1e-3 * 1024
    grads = torch.zeros_like(flat)
    buf = torch.zeros(64, dtype=torch.int64, device=DEV)
    for _ in range(2):
        ops.mlp_backward_pipe(grads, up, out, stash, ws, packed, flat, n_rays, s, 1024.0)
    buf.zero_()
    _lib.load().nerf_debug_set_timing(buf.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.mlp_backward_pipe(grads, up, out, stash, ws, packed, flat, n_rays, s, 1024.0)
    b.record()
    torch.cuda.synchronize()
    _lib.load().nerf_debug_set_timing(None)
    v = buf.tolist()
    n_stage, n_head = max(v[19], 1), max(v[20], 1)
    print(f'--- fused backward {n_rays}x{s}: {a.elapsed_time(b):.3f} ms (pipeline + residual wgrad); {v[19]} stage leaders, {v[20]} head slot threads')
    for i, name in NAMES.items():
        denom = n_head if i >= 13 else n_stage
        print(f'    {name:28s} {v[i] / denom:12.0f} cyc')
    print('    per role (8 pipelines each): [loader wait ready flag | mma wait operands | store busy] in K cycles per group')
    groups = max(1, ((n + 127) // 128 + 1) // 2 // 8)
    for role in range(9):
        r = [v[32 + 3 * role + m] / 8 / groups / 1e3 for m in range(3)]
        print(f'      role {role} ({"head" if role == 0 else "stage " + str(role)}): ' + '  '.join(f'{x:7.2f}' for x in r) + ('   (head: col 0 = wait slot freed, summed over 32 slot threads)' if role == 0 else ''))
    a.record()
    ops.mlp_backward_legacy(grads, up, out, stash, ws, packed, flat, n_rays, s, 1024.0)
    b.record()
    torch.cuda.synchronize()
    print(f'--- legacy dgrad + wgrad: {a.elapsed_time(b):.3f} ms')


if __name__ == '__main__':
    main()
