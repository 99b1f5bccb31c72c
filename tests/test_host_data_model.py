"""CPU: host-side data model and loop shell of the plugin (RayBatch / RayCollection, cameras, samplers, NeRFLoss against
the oracle, callback ordering of the base trainer).  No compute kernels are involved."""
import math

import pytest
import torch

from nerficg_b200 import Framework
from oracle import nerf_oracle as O


@pytest.fixture()
def cfg():
    Framework.load_config(None, {'RENDERER.N_SAMPLES': 192, 'RENDERER.COARSE_RATIO': 1 / 3 + 1e-7, 'GLOBAL.LOG_LEVEL': 0})
    yield Framework.config


def _batch(n, g, with_alpha=True):
    from nerficg_b200.Datasets import RayBatch
    mk = lambda c: torch.rand(n, c, generator=g)
    return RayBatch(origin=mk(3), direction=mk(3), view_direction=mk(3), rgb=mk(3), alpha=mk(1) if with_alpha else None)


def test_ray_batch_semantics(cfg):
    """Field validation, slicing, index tensors, split / cat (reference Datasets/utils.py:536-670)."""
    from nerficg_b200.Datasets import RayBatch, RayCollection
    g = torch.Generator().manual_seed(0)
    b = _batch(10, g)
    assert len(b) == 10 and b.dtype == torch.float32 and b.device.type == 'cpu'
    assert b[...] is b and b[:] is b
    assert len(b[3]) == 1 and torch.equal(b[3].origin, b.origin[3:4])
    ids = torch.tensor([7, 0, 7])
    assert torch.equal(b[ids].rgb, b.rgb[ids]) and b[ids].depth is None          # CPU index tensors take the torch path
    parts = b.split(4)
    assert [len(p) for p in parts] == [4, 4, 2]
    assert torch.equal(RayBatch.cat(parts).direction, b.direction)
    with pytest.raises(Framework.DatasetError):
        RayBatch(origin=torch.zeros(4, 3), direction=torch.zeros(5, 3))
    with pytest.raises(Framework.DatasetError):
        RayBatch(origin=torch.zeros(4, 3), direction=torch.zeros(4, 3, dtype=torch.float64))
    with pytest.raises(Framework.DatasetError):
        RayBatch.cat([b, _batch(3, g, with_alpha=False)])                        # a field present in only some batches
    with pytest.raises(Framework.DatasetError):
        RayBatch.cat([])
    col = RayCollection(b, [slice(0, 6), slice(6, 10)])
    assert len(col) == 10 and len(col[1]) == 4 and col.all_rays is b


def test_camera_helpers_and_local_directions(cfg):
    from nerficg_b200.Cameras import PerspectiveCamera, SharedCameraSettings, focal_to_fov, fov_to_focal
    assert fov_to_focal(0.6911112070083618) * 800 == pytest.approx(1111.111, abs=1e-3)
    assert focal_to_fov(fov_to_focal(0.7)) == pytest.approx(0.7)
    assert fov_to_focal(90.0, degrees=True) == pytest.approx(0.5)
    with pytest.raises(Framework.CameraError):
        SharedCameraSettings(torch.ones(4), 2.0, 6.0)
    with pytest.raises(Framework.CameraError):
        SharedCameraSettings(torch.ones(3), 6.0, 2.0)
    shared = SharedCameraSettings(torch.ones(3), 2.0, 6.0)
    cam = PerspectiveCamera(shared_settings=shared, width=7, height=5, focal_x=9.0, center_x=3.1)
    assert cam.focal_y == 9.0 and cam.center_y == 2.5 and (cam.near_plane, cam.far_plane) == (2.0, 6.0)
    local = cam.compute_local_ray_directions(device='cpu')
    ref = O.camera_rays(torch.eye(4, dtype=torch.float64), 7, 5, 9.0, 9.0, 3.1, 2.5)[1]   # identity pose: world == camera axes
    assert local.shape == (35, 3) and (local - ref).abs().max() <= 1e-6
    assert cam.compute_local_ray_directions(device='cpu') is local                         # cached per parameter set
    cam.background_color = torch.zeros(3)
    assert torch.equal(shared.background_color, torch.zeros(3))                            # shared between all cameras


def test_samplers(cfg):
    from nerficg_b200.Optim.Samplers import DatasetSampler, RandomImageSampler, RandomSequentialSampler, RayPoolSampler
    torch.manual_seed(0)
    ids = RandomImageSampler(50).get(1000)
    assert ids.dtype == torch.int64 and int(ids.min()) >= 0 and int(ids.max()) < 50 and len(ids.unique()) > 40
    seq = RandomSequentialSampler(5)
    first = sorted(int(seq.get(1)) for _ in range(5))
    second = sorted(int(seq.get(1)) for _ in range(5))
    assert first == second == [0, 1, 2, 3, 4]                                              # every element once per epoch

    class FakeDataset:
        mode = 'train'

        def __init__(self):
            from nerficg_b200.Cameras import PerspectiveCamera, SharedCameraSettings
            cam = PerspectiveCamera(shared_settings=SharedCameraSettings(torch.ones(3), 2.0, 6.0), width=4, height=3)
            self.views = [type('V', (), {'camera': cam})() for _ in range(3)]

        def __len__(self):
            return len(self.views)

        def __iter__(self):
            return iter(self.views)

        def __getitem__(self, i):
            return self.views[i]

        def get_total_ray_count(self):
            return 36

    ds = FakeDataset()
    sampler = DatasetSampler(ds)
    assert len(sampler.img_samplers) == 3 and sampler.img_samplers[0].num_elements == 12
    info = sampler.get(ds, ray_batch_size=None)                                            # view only, no rays requested
    assert info['ray_batch'] is None and info['view'] is ds[info['sample_id']]
    ds.mode = 'test'
    with pytest.raises(Framework.SamplerError):
        sampler.get(ds, ray_batch_size=None)
    with pytest.raises(Framework.SamplerError):
        RayPoolSampler(FakeDataset()).get(ds, 8)


@pytest.mark.parametrize('lam_a,coarse', [(0.0, True), (0.4, True), (0.4, False)])
def test_nerf_loss_module_matches_oracle(cfg, lam_a, coarse):
    """The torch loss module of the autograd path == the oracle's restatement of the reference (Loss.py:26-43)."""
    from nerficg_b200.Methods.NeRF.Loss import NeRFLoss
    g = torch.Generator().manual_seed(3)
    n = 257
    rays = _batch(n, g)
    out = {'rgb': torch.rand(n, 3, generator=g), 'alpha': torch.rand(n, 1, generator=g)}
    if coarse:
        out |= {'rgb_coarse': torch.rand(n, 3, generator=g), 'alpha_coarse': torch.rand(n, 1, generator=g)}
    bg = torch.tensor([1.0, 0.5, 0.0])
    loss = NeRFLoss(0.7, lam_a, coarse)(out, rays, bg)
    ref = O.nerf_loss(out, rays.rgb, rays.alpha, bg, lambda_color=0.7, lambda_alpha=lam_a)
    assert float(loss) == pytest.approx(float(ref), rel=1e-6)
    from nerficg_b200.Methods.NeRF.Loss import peak_signal_noise_ratio
    a, b = torch.rand(3, 8, 8, generator=g), torch.rand(3, 8, 8, generator=g)
    assert float(peak_signal_noise_ratio(a, b)) == pytest.approx(O.psnr(a, b), rel=1e-6)
    from nerficg_b200.Optim.Losses import BaseLoss
    container = BaseLoss()
    container.add_loss_metric('X', torch.nn.functional.mse_loss, 1.0)
    container.add_loss_metric('Off', torch.nn.functional.mse_loss, 0.0)          # weight 0: constant 0, never evaluated
    assert float(container({'X': {'input': a, 'target': b}, 'Off': {}})) == pytest.approx(float(torch.mean((a - b) ** 2)))
    with pytest.raises(Framework.LossError):
        container({'Off': {}})                                                   # no argument configuration for 'X'


def test_trainer_callbacks_run_in_priority_order(cfg):
    """Callback discovery, ordering, stride / start / end windows of the base loop (reference Base/Trainer.py:225-291)."""
    from nerficg_b200.Methods.Base.Trainer import BaseTrainer
    from nerficg_b200.Methods.Base.utils import post_training_callback, pre_training_callback, training_callback
    from nerficg_b200.Methods.NeRF.Model import NeRF
    from nerficg_b200.Methods.NeRF.Renderer import NeRFRenderer
    Framework.config.TRAINING.NUM_ITERATIONS = 6
    log = []

    class T(BaseTrainer):
        @pre_training_callback(priority=10)
        def pre_low(self, it, ds):
            log.append(('pre_low', it))

        @pre_training_callback(priority=1000)
        def pre_high(self, it, ds):
            log.append(('pre_high', it))

        @training_callback(priority=50)
        def step(self, it, ds):
            log.append(('step', it))

        @training_callback(priority=100, start_iteration=2, end_iteration=4, iteration_stride=2)
        def windowed(self, it, ds):
            log.append(('windowed', it))

        @training_callback(active='RUN_VALIDATION', priority=60)
        def validation(self, it, ds):
            log.append(('validation', it))

        @training_callback(active=True, iteration_stride=-1)
        def disabled_by_stride(self, it, ds):
            log.append(('disabled', it))

        @post_training_callback(priority=1)
        def post(self, it, ds):
            log.append(('post', it))

    model = NeRF('t').build()
    trainer = T(model=model, renderer=NeRFRenderer(model))
    trainer.run(dataset=None)
    assert log[:2] == [('pre_high', 0), ('pre_low', 0)]
    assert [it for name, it in log if name == 'step'] == [0, 1, 2, 3, 4, 5]
    assert [it for name, it in log if name == 'windowed'] == [2, 4]
    assert log.index(('windowed', 2)) < log.index(('step', 2))                 # higher priority first within an iteration
    assert not any(name in ('validation', 'disabled') for name, _ in log)      # RUN_VALIDATION is off; stride <= 0 disables
    assert log[-1] == ('post', 6) and model.num_iterations_trained == 6

    class Bad(BaseTrainer):
        @training_callback(active='NO_SUCH_KEY')
        def step(self, it, ds):
            pass
    with pytest.raises(Framework.TrainingError):
        Bad(model=model, renderer=NeRFRenderer(model)).run(dataset=None)
