"""CPU: host-side mirror of the reference's plugin interface (config, registry, model layout, schedule)."""
import math

import pytest
import torch

from nerficg_b200 import Framework, dist, params


@pytest.fixture()
def cfg():
    Framework.load_config(None, {'RENDERER.N_SAMPLES': 192, 'RENDERER.COARSE_RATIO': 1 / 3 + 1e-7, 'GLOBAL.LOG_LEVEL': 0})
    yield Framework.config


def test_config_overrides_and_errors(cfg):
    assert cfg.RENDERER.N_SAMPLES == 192
    Framework.load_config(None, {'TRAINING.BATCH_SIZE': '4096', 'GLOBAL.LOG_LEVEL': 0})
    assert Framework.config.TRAINING.BATCH_SIZE == 4096          # strings are literal_eval'ed like KEY=VAL overrides
    with pytest.raises(Framework.FrameworkError):
        Framework.load_config(None, {'NOPE.X.Y': 1, 'GLOBAL.LOG_LEVEL': 0})


def test_plugin_triple_and_registry(cfg):
    from nerficg_b200.Implementations import Methods, install_into_reference
    mod = Methods.import_method('NeRF')
    assert {mod.MODEL.__name__, mod.RENDERER.__name__, mod.TRAINING_INSTANCE.__name__} == {'NeRF', 'NeRFRenderer', 'NeRFTrainer'}
    with pytest.raises(Framework.MethodError):
        Methods.import_method('InstantNGP')

    class FakeFramework:            # stands for the reference's Framework module (the real one: tests/test_reference_dropin.py)
        class Configurable: pass
        class Directories: OUTPUT_DIR = NERFICG_ROOT = None
        config = Framework.ConfigParameterList.fromDict({'GLOBAL': {'METHOD_TYPE': 'NeRF'}, 'RENDERER': {'N_SAMPLES': 48, 'COARSE_RATIO': 0.5},
                                                         'TRAINING': {'WANDB': {'ACTIVATE': False}}})

    class FakeReference:            # stands for the reference's Implementations module
        Framework = FakeFramework
        class Methods:
            modules = {}
    from nerficg_b200.Implementations import uninstall_from_reference
    install_into_reference(FakeReference)
    try:
        assert FakeReference.Methods.modules['NeRF'] is mod
        assert Framework.config is FakeFramework.config and Framework.Directories is FakeFramework.Directories
        FakeFramework.config = Framework.ConfigParameterList.fromDict({'GLOBAL': {}, 'RENDERER': {'N_SAMPLES': 12}})   # host rebinds its global
        assert Framework.config.RENDERER.N_SAMPLES == 12
        with pytest.raises(Framework.FrameworkError):
            Framework.load_config(None)          # while bound, the configuration belongs to the host
    finally:
        uninstall_from_reference(FakeReference)
    assert Framework.config is not FakeFramework.config


def test_model_state_dict_matches_reference_layout(cfg):
    from nerficg_b200.Methods.NeRF.Model import NeRF
    model = NeRF('t').build()
    sd = model.state_dict()
    for prefix in ('coarse_nerf.', 'nerf.'):
        for name, shape in params.TENSOR_SPECS:
            assert tuple(sd[prefix + name].shape) == shape
        assert tuple(sd[prefix + 'encoding_position.frequency_factors'].shape) == (1, 1, 10)
        assert tuple(sd[prefix + 'encoding_direction.frequency_factors'].shape) == (1, 1, 4)
    assert sum(p.numel() for p in model.parameters()) == 1191688
    # parameters are views into one flat buffer per block, in C-layout order
    block = model.nerf
    flat = block.flat_params
    off, size, _ = params.layout()
    for p, o in zip(block.ordered_parameters(), off):
        assert p.data_ptr() == flat.data_ptr() + 4 * o
    w = block.initial_layers[5][0].weight
    assert tuple(w.shape) == (256, 319)                          # skip concat (h, enc) before layer 5
    # same init distribution as nn.Linear default: U(+-1/sqrt(fan_in))
    assert w.abs().max().item() <= 1 / math.sqrt(319) + 1e-7


def test_unsupported_architecture_is_rejected(cfg):
    from nerficg_b200.Methods.NeRF.Model import NeRFBlock
    with pytest.raises(Framework.ModelError):
        NeRFBlock(8, 1, 128, 10, 4, True, [5], 'relu')
    with pytest.raises(Framework.ModelError):
        NeRFBlock(8, 1, 256, 10, 4, True, [5], 'softplus')


def test_renderer_sample_split_and_model_check(cfg):
    from nerficg_b200.Methods.NeRF.Model import NeRF
    from nerficg_b200.Methods.NeRF.Renderer import NeRFRenderer, default_grad_scale
    r = NeRFRenderer(NeRF('t').build())
    assert (r.n_samples_coarse_nerf, r.n_samples_nerf) == (64, 128)   # round(N_SAMPLES * COARSE_RATIO), Renderer.py:113-114
    assert r.RAY_BATCH_SIZE == 8192
    with pytest.raises(Framework.RendererError):
        NeRFRenderer(torch.nn.Linear(1, 1))
    assert default_grad_scale(4096) == 65536.0


def test_lr_schedule_is_log_linear():
    from nerficg_b200.Optim.lr_utils import LRDecayPolicy
    f = LRDecayPolicy(lr_init=5e-4, lr_final=5e-5, max_steps=1000)
    assert f(0) == pytest.approx(5e-4) and f(1000) == pytest.approx(5e-5) and f(5000) == pytest.approx(5e-5)
    assert f(500) == pytest.approx(math.sqrt(5e-4 * 5e-5))


def test_shard_range_partitions_exactly():
    for n, w in ((200, 8), (200, 3), (5, 8), (0, 4), (640000, 7)):
        seen = []
        for r in range(w):
            seen += list(dist.shard_range(n, r, w))
        assert seen == list(range(n))
        sizes = [len(dist.shard_range(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dist.shard_range(10, 4, 4)


def test_depth_color_map_matches_reference(golden):
    """apply_color_map / SPECTRAL against the reference's own output (oracle/make_golden_postprocess.py)."""
    from nerficg_b200.Visual import apply_color_map, spectral_lut
    g = golden('postprocess')
    lut = spectral_lut()
    assert lut.shape == (256, 3) and float(lut.min()) >= 0.0 and float(lut.max()) <= 1.0
    nf = (g['near'], g['far'])
    assert (apply_color_map('SPECTRAL', g['depth'], nf, g['alpha']) - g['spectral_masked']).abs().max() <= 1e-6
    assert (apply_color_map('SPECTRAL', g['depth'], None, g['alpha']) - g['spectral_auto']).abs().max() <= 1e-6
    assert (apply_color_map('SPECTRAL', g['depth'], nf) - g['spectral_plain']).abs().max() <= 1e-6
    assert (apply_color_map('Grayscale', g['depth'], nf, invert=True) - g['gray_inverted']).abs().max() <= 1e-6
    import pytest
    from nerficg_b200 import Framework
    with pytest.raises(Framework.RendererError):
        apply_color_map('SPECTRAL', g['depth'][0], nf)


def test_postprocess_outputs_shapes_and_depth_coloring(cfg, golden):
    """NeRFRenderer.postprocess_outputs (reference Renderer.py:142-165): six 3xHxW images in [0, 1]; depth = SPECTRAL over
    [near, far] masked by alpha."""
    from nerficg_b200.Cameras import PerspectiveCamera, SharedCameraSettings
    from nerficg_b200.Datasets import View
    from nerficg_b200.Methods.NeRF.Model import NeRF
    from nerficg_b200.Methods.NeRF.Renderer import NeRFRenderer
    import numpy as np
    g = golden('postprocess')
    h, w = g['depth'].shape[1:]
    cam = PerspectiveCamera(shared_settings=SharedCameraSettings(torch.ones(3), g['near'], g['far']), width=w, height=h)
    view = View(cam, np.eye(4))
    r = NeRFRenderer(NeRF('t').build())
    gen = torch.Generator().manual_seed(1)
    rgb = torch.rand(3, h, w, generator=gen) * 1.4 - 0.2
    outputs = {'rgb': rgb.clone(), 'alpha': g['alpha'].clone(), 'depth': g['depth'].clone(),
               'rgb_coarse': rgb.clone(), 'alpha_coarse': g['alpha'].clone(), 'depth_coarse': g['depth'].clone()}
    out = r.postprocess_outputs(outputs, view, None, 0)
    assert set(out) == {'rgb', 'alpha', 'depth', 'rgb_coarse', 'alpha_coarse', 'depth_coarse'}
    for k, v in out.items():
        assert tuple(v.shape) == (3, h, w) and float(v.min()) >= 0.0 and float(v.max()) <= 1.0, k
    assert torch.equal(out['rgb'], rgb.clamp(0, 1))
    assert (out['depth'] - g['spectral_masked']).abs().max() <= 1e-6
    assert (out['depth_coarse'] - g['spectral_masked']).abs().max() <= 1e-6


def test_ssim_matches_an_independent_implementation():
    """SSIM of the test-set metrics (reference Base/Renderer.py:130-133, torchmetrics defaults restated): identical images
    give 1, the measure is symmetric, and it equals a scipy Gaussian-filter implementation on the un-padded interior."""
    import numpy as np
    from scipy.ndimage import correlate
    from nerficg_b200.Methods.Base.Renderer import psnr_8bit, ssim
    g = torch.Generator().manual_seed(0)
    a = torch.rand(3, 40, 50, generator=g)
    b = (a + 0.1 * torch.randn(3, 40, 50, generator=g)).clamp(0, 1)
    assert ssim(a, a) == pytest.approx(1.0, abs=1e-6) and ssim(a, b) == pytest.approx(ssim(b, a), abs=1e-7)
    x = np.arange(11) - 5
    g1 = np.exp(-(x / 1.5) ** 2 / 2)
    k = np.outer(g1 / g1.sum(), g1 / g1.sum())
    f = lambda img: np.stack([correlate(c, k, mode='mirror') for c in img])
    A, B = a.numpy().astype(np.float64), b.numpy().astype(np.float64)
    mu_a, mu_b = f(A), f(B)
    va, vb, cab = f(A * A) - mu_a ** 2, f(B * B) - mu_b ** 2, f(A * B) - mu_a * mu_b
    m = ((2 * mu_a * mu_b + 1e-4) * (2 * cab + 9e-4)) / ((mu_a ** 2 + mu_b ** 2 + 1e-4) * (va + vb + 9e-4))
    assert ssim(a, b) == pytest.approx(m[:, 5:-5, 5:-5].mean(), abs=1e-5)
    assert psnr_8bit(a, a) > 100.0
