"""Tensor-pipe rate of cta_group::2 M256 N256 K16 MMAs with K-major vs MN-major operands (nerf_selftest_umma2_rate)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nerficg_b200 import _lib  # noqa: E402
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device='cuda:0')
for mn, name in ((0, 'K-major'), (1, 'MN-major')):
    for iters in (64, 1024, 4096):
        for _ in range(2):
            out.zero_()
            _lib.check(lib.nerf_selftest_umma2_rate(out.data_ptr(), mn, iters, _lib.stream_ptr()), 'umma2_rate')
            torch.cuda.synchronize()
        print(f'{name:9s} {iters:5d} MMAs: {out[0].item() / iters:7.1f} cycles per MMA (floor 128)')
