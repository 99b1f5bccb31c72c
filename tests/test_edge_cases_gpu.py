"""Edge cases of the hot path: empty batches, a single ray, the smallest sample counts, empty space
(all-zero coarse weights), ray counts around the tile and CTA-pair boundaries.  Everything is checked against the fp32
oracle (oracle/nerf_oracle.py), which follows the reference's own code for these cases."""
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def ops():
    from nerficg_b200 import ops
    return ops


@pytest.fixture(scope='module')
def fw():
    from nerficg_b200 import Framework
    Framework.setup(None, {'RENDERER.N_SAMPLES': 192, 'RENDERER.COARSE_RATIO': 0.3333333, 'RENDERER.RAY_BATCH_SIZE': 16,
                           'GLOBAL.LOG_LEVEL': 0})
    return Framework


def _rays(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g) * 0.1 + torch.tensor([0.0, -4.0, 0.5])
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.2 + torch.tensor([0.0, 1.0, -0.1]), dim=-1) * 1.05
    v = torch.nn.functional.normalize(d, dim=-1)
    return o.to(DEV), d.to(DEV), v.to(DEV)


def test_empty_batch_is_a_no_op_everywhere(ops, fw):
    """Zero rays: every C entry returns success without a launch; the renderer raises like the reference does."""
    from nerficg_b200 import params
    from nerficg_b200.Cameras import PerspectiveCamera, SharedCameraSettings
    from nerficg_b200.Datasets import RayBatch
    from nerficg_b200.Implementations import Methods
    dev = torch.device(DEV)
    z = ops.sample_stratified(0, 64, 2.0, 6.0, None, dev)
    assert z.shape == (0, 64)
    merged = ops.sample_importance(z, torch.empty(0, 64, device=DEV), 128, None)
    merged = merged[0] if isinstance(merged, tuple) else merged
    assert merged.shape == (0, 192)
    flat = torch.zeros(params.layout()[2], device=DEV)
    o, d, v = _rays(0)
    rs = ops.mlp_forward(ops.mlp_pack(flat, with_backward=False), flat, o, d, v, z)
    assert rs.shape == (0, 64, 4)
    rgb, depth, alpha, _ = ops.composite_forward(z, rs, d, torch.ones(3, device=DEV))
    assert rgb.shape == (0, 3) and depth.numel() == 0 and alpha.numel() == 0
    model = Methods.get_model('NeRF', name='t')
    renderer = Methods.get_renderer('NeRF', model)
    cam = PerspectiveCamera(shared_settings=SharedCameraSettings(torch.ones(3), 2.0, 6.0), width=10, height=10, focal_x=10.0, focal_y=10.0)
    # the renderer mirrors the reference's error behaviour: an empty RayBatch splits into zero chunks and the reference's
    # `outputs[key][0]` (src/Methods/NeRF/Renderer.py:94) raises IndexError -- checked against the live reference on CPU
    with torch.no_grad(), pytest.raises(IndexError):
        renderer.render_rays(RayBatch(origin=o, direction=d, view_direction=v), cam)


@pytest.mark.parametrize('n_samples_total', [1, 127, 128, 129, 255, 256, 257, 511, 512, 513, 148 * 256 - 1, 148 * 256 + 1])
def test_sample_counts_around_tile_pair_and_wave_boundaries(ops, n_samples_total):
    """One tile is 128 samples, a CTA pair takes two tiles per slot, one wave of the persistent grid is 148 x 256 samples:
    totals that straddle those edges (one sample per ray, so the last tile is ragged) against the fp32 oracle, inference and
    training variants, plus the backward's handling of the ragged tail (finite gradients, zero where no sample contributes)."""
    from nerficg_b200 import params
    sd = O.init_state_dict(0)
    flat = params.flatten_state_dict(sd, 'nerf.', DEV)
    packed = ops.mlp_pack(flat)
    n = n_samples_total
    g = torch.Generator().manual_seed(n)
    o = torch.randn(n, 3, generator=g) * 2
    d = torch.randn(n, 3, generator=g)
    vd = torch.nn.functional.normalize(d, dim=-1)
    z = 2 + 4 * torch.rand(n, 1, generator=g)
    x = o + d * z
    sig, rgb = O.mlp_forward(sd, 'nerf.', x, vd, None)
    out = ops.mlp_forward(packed, flat, o.to(DEV), d.to(DEV), vd.to(DEV), z.to(DEV)).reshape(-1, 4)
    assert (out[:, :3].cpu() - rgb).abs().max() <= 1e-3
    assert (out[:, 3:].cpu() - sig).abs().max() <= 1e-3 + 2e-3 * sig.abs().max()
    stash = torch.empty(ops.mlp_stash_bytes(n), dtype=torch.uint8, device=DEV)
    out_t = ops.mlp_forward(packed, flat, o.to(DEV), d.to(DEV), vd.to(DEV), z.to(DEV), None, stash)
    assert torch.equal(out_t.reshape(-1, 4), out)
    grads = torch.zeros_like(flat)
    ws = torch.empty(ops.mlp_backward_workspace_bytes(n), dtype=torch.uint8, device=DEV)
    up = torch.zeros(n, 4, device=DEV)
    ops.mlp_backward(grads, up, out_t, stash, ws, packed, flat, n, 1, 1.0)
    assert bool((grads == 0).all())                      # zero upstream gradient -> exactly zero parameter gradient (no stale tail rows)
    up = torch.randn(n, 4, generator=g).to(DEV) * 1e-3
    ops.mlp_backward(grads, up, out_t, stash, ws, packed, flat, n, 1, 1.0)
    assert bool(torch.isfinite(grads).all()) and float(grads.abs().max()) > 0.0


def test_importance_sampling_in_empty_space(ops):
    """All-zero coarse weights (a ray that hits nothing): the reference's pdf is the uniform 1e-5 floor (utils.py:82-84), so
    the fine depths are a uniform resampling of the coarse bins; also a single spike and weights that underflow in the sum."""
    n, nc, nf = 64, 64, 128
    g = torch.Generator().manual_seed(11)
    z = torch.sort(2 + 4 * torch.rand(n, nc, generator=g), -1).values
    w = torch.zeros(n, nc)
    w[1, 17] = 1.0                                  # one opaque sample
    w[2] = 1e-30                                    # far below the 1e-5 floor
    w[3, 1:-1] = torch.rand(nc - 2, generator=g)    # the generic case next to them
    u = torch.rand(n, nf, generator=g)
    ref_fine = O.importance_depths(z, w, nf, u)
    ref = O.merge_depths(z, ref_fine)
    got = ops.sample_importance(z.to(DEV), w.to(DEV), nf, u.to(DEV))
    got = got[0] if isinstance(got, tuple) else got
    assert got.shape == ref.shape
    assert bool((got[:, 1:] >= got[:, :-1]).all())
    # same criterion as tests/test_stages_gpu.py: positions agree to 2e-5 except where an ulp of the cdf moves a sample across a
    # bracket whose span is below 1e-5 (the reference's `denom < 1e-5 -> 1` branch is discontinuous there)
    err = (got.cpu() - ref).abs()
    assert (err > 2e-5).float().mean() <= 5e-3 and err.max() <= 0.1, ((err > 2e-5).float().mean(), err.max())
    assert (err[0] <= 2e-5).all() and (err[2] <= 2e-5).all()      # the two uniform-pdf rays have no such bracket


def test_single_ray_full_render(fw):
    """One ray through the whole renderer (stratified -> coarse -> importance -> fine -> compositing) against the oracle."""
    from nerficg_b200.Cameras import PerspectiveCamera, SharedCameraSettings
    from nerficg_b200.Datasets import RayBatch
    from nerficg_b200.Implementations import Methods
    sd = O.init_state_dict(5)
    model = Methods.get_model('NeRF', name='t')
    model.load_state_dict(sd, strict=True)
    renderer = Methods.get_renderer('NeRF', model)
    bg = torch.tensor([1.0, 0.5, 0.25])
    cam = PerspectiveCamera(shared_settings=SharedCameraSettings(bg, 2.0, 6.0), width=10, height=10, focal_x=10.0, focal_y=10.0)
    o, d, v = _rays(1, seed=9)
    with torch.no_grad():
        got = renderer.render_rays(RayBatch(origin=o, direction=d, view_direction=v), cam)
        ref = O.render_rays(sd, o.cpu(), d.cpu(), v.cpu(), 2.0, 6.0, bg, 64, 128)
    for k in ('rgb', 'rgb_coarse', 'alpha', 'alpha_coarse'):
        assert (got[k].cpu().reshape(-1) - ref[k].reshape(-1)).abs().max() <= 1e-3, k
