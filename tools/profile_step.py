"""One eager training iteration at the bench configuration (for ncu): every C-ABI kernel is launched `reps` times."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from nerficg_b200 import Framework  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Framework.setup(None, {'RENDERER.N_SAMPLES': 192, 'RENDERER.COARSE_RATIO': 1 / 3 + 1e-7, 'TRAINING.BATCH_SIZE': 4096,
                       'TRAINING.NUM_ITERATIONS': 500000, 'GLOBAL.LOG_LEVEL': 0})
from nerficg_b200.Datasets.Synthetic import SyntheticLegoDataset  # noqa: E402
from nerficg_b200.Implementations import Methods  # noqa: E402
from nerficg_b200.Methods.NeRF.Trainer import _FusedStep  # noqa: E402
torch.manual_seed(0)
trainer = Methods.get_training_instance('NeRF')
dev = Framework.config.GLOBAL.DEFAULT_DEVICE
ds = SyntheticLegoDataset(200, 200, 2, 1, device=dev)
ds.precompute_rays(['train'])
pool = ds.ray_collection['train'].all_rays
step = _FusedStep(trainer, 4096, ds.default_camera, False)
batch = pool[torch.randint(0, len(pool), (4096,), device=dev)]
step.origin.copy_(batch.origin); step.direction.copy_(batch.direction); step.view_direction.copy_(batch.view_direction)
step.rgb_gt.copy_(batch.rgb); step.alpha_gt.copy_(batch.alpha)
times = bench.instrumented_kernel_times(trainer, step, reps=reps)
print({k: round(v, 4) for k, v in times.items()})
