"""TMEM read bandwidth of one SM (nerf_selftest_tmem_read): sizes the epilogue floor of the MLP chain kernels."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nerficg_b200 import _lib  # noqa: E402

lib = _lib.load()
out = torch.zeros(4, dtype=torch.int64, device='cuda:0')
for mode, name in ((0, 'x32'), (1, '2 x x32 in flight'), (2, 'x16')):
    for warps in (1, 4, 8, 16):
        for _ in range(2):
            _lib.check(lib.nerf_selftest_tmem_read(out.data_ptr(), warps, mode, 2000, _lib.stream_ptr()), 'tmem probe')
        torch.cuda.synchronize()
        cyc, nbytes = out[0].item(), out[1].item()
        print(f'{name:18s} warps={warps:2d}: {cyc:8d} cycles, {nbytes / cyc:7.1f} B/cycle/SM, {cyc / 2000:6.1f} cycles per load round')
