"""L2 -> SM streaming bandwidth through the TMA engines (development aid; nerf_selftest_l2_stream): bytes per cycle and SM for
loads, stores and both, from an L2-resident window, with 148 / 74 / 16 CTAs."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nerficg_b200 import _lib  # noqa: E402

DEV = 'cuda:0'
lib = _lib.load()
win = torch.zeros(32 << 20, dtype=torch.uint8, device=DEV)
out = torch.zeros(2, dtype=torch.int64, device=DEV)
for n_ctas in (148, 74, 16):
    for mode, name in ((1, 'loads'), (2, 'stores'), (3, 'loads+stores')):
        for rep in range(2):
            out.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.nerf_selftest_l2_stream(out.data_ptr(), win.data_ptr(), win.numel(), mode, 4096, n_ctas, _lib.stream_ptr()), 'l2_stream')
            b.record()
            torch.cuda.synchronize()
        cyc, nbytes = out.tolist()
        ms = a.elapsed_time(b)
        print(f'{n_ctas:4d} CTAs {name:13s}: {nbytes / cyc:7.1f} B/clk/SM  {nbytes * n_ctas / cyc:8.0f} B/clk chip  {nbytes * n_ctas / ms / 1e9:7.2f} TB/s  ({ms:.3f} ms)')

print('--- TMA bulk loads, chunk size sweep (148 CTAs, 128 KB ring per CTA)')
for csel, kb in ((1, 2), (2, 4), (3, 8), (0, 16), (5, 32), (6, 64)):
    iters = 4096 * 16 // kb
    for rep in range(2):
        out.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.nerf_selftest_l2_stream(out.data_ptr(), win.data_ptr(), win.numel(), 1 | (csel << 4), iters, 148, _lib.stream_ptr()), 'l2_stream')
        b.record()
        torch.cuda.synchronize()
    cyc, nbytes = out.tolist()
    print(f'  chunk {kb:3d} KB: {nbytes / cyc:7.1f} B/clk/SM   {cyc / iters:7.1f} clk per op')
import sys
if '--chunks-only' in sys.argv:
    sys.exit(0)
out3 = torch.zeros(4, dtype=torch.int64, device=DEV)
for n_ctas in (148, 16):
    for n_warps in (2, 4, 8):
        for mode, name in ((4, 'LSU stores'), (8, 'LSU loads'), (5, 'TMA loads + LSU stores'), (1, 'TMA loads alone (lsu kernel)')):
            for rep in range(2):
                out3.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _lib.check(lib.nerf_selftest_l2_stream_lsu(out3.data_ptr(), win.data_ptr(), win.numel(), mode, 2048, n_ctas, n_warps, _lib.stream_ptr()), 'l2_stream_lsu')
                b.record()
                torch.cuda.synchronize()
            cyc, nbytes = out3.tolist()[:2]
            print(f'{n_ctas:4d} CTAs {n_warps} warps {name:28s}: {nbytes / cyc:7.1f} B/clk/SM (sum of directions)  {nbytes * n_ctas / a.elapsed_time(b) / 1e9:7.2f} TB/s')
