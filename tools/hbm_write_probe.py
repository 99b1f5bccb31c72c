"""Pure-write / read / read+write HBM bandwidth with torch kernels (development aid): are the stash-writing chain kernels
(4.1-4.4 TB/s of stores) bound by HBM writes?  B200: fp32 fill 7.5 TB/s (compressible data), in-place mul 6.9 TB/s, sum 6.45 TB/s."""
import torch
x = torch.empty(8 << 30, dtype=torch.uint8, device='cuda')
y = torch.empty(1 << 30, dtype=torch.float32, device='cuda')
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.fill_(1)); print(f'memset-style fill 8 GiB: {ms:.3f} ms = {(8<<30)/ms/1e9:.2f} TB/s write')
ms = t(lambda: y.fill_(1.5)); print(f'fp32 fill 4 GiB: {ms:.3f} ms = {(4<<30)/ms/1e9:.2f} TB/s write')
ms = t(lambda: y.mul_(1.0001)); print(f'in-place mul 4 GiB: {ms:.3f} ms = {(8<<30)/ms/1e9:.2f} TB/s read+write')
ms = t(lambda: y.sum()); print(f'sum 4 GiB: {ms:.3f} ms = {(4<<30)/ms/1e9:.2f} TB/s read')
