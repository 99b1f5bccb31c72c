"""One 800 x 800 view through NeRFRenderer.render_image at the bench's render configuration (config C), bracketed by
cudaProfilerStart/Stop for `ncu --profile-from-start off --metrics gpu__time_duration.sum`; also prints the CUDA-event time of
the view and of an (K0-free) render_rays call so that launch gaps can be told from kernel time."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nerficg_b200 import Framework  # noqa: E402

Framework.setup(None, {'RENDERER.N_SAMPLES': 192, 'RENDERER.COARSE_RATIO': 1 / 3 + 1e-7, 'GLOBAL.LOG_LEVEL': 0})
from nerficg_b200.Datasets.Synthetic import SyntheticLegoDataset  # noqa: E402
from nerficg_b200.Implementations import Methods  # noqa: E402
torch.manual_seed(0)
dev = Framework.config.GLOBAL.DEFAULT_DEVICE
model = Methods.get_model('NeRF', name='p')
renderer = Methods.get_renderer('NeRF', model)
renderer.RAY_BATCH_SIZE = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ds = SyntheticLegoDataset(800, 800, 1, 2, device=dev)
views = ds.test()
model.eval()
with torch.no_grad():
    renderer.render_image(views[0])
    torch.cuda.synchronize()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    out = renderer.render_image(views[1])
    b.record()
    rays = views[1].get_rays()
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    renderer.render_rays(rays, views[1].camera)
    c1.record()
    torch.cuda.synchronize()
    print(f'render_image {a.elapsed_time(b):.2f} ms, render_rays alone {c0.elapsed_time(c1):.2f} ms, chunk {renderer.RAY_BATCH_SIZE}')
    torch.cuda.profiler.start()
    renderer.render_image(views[0])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
