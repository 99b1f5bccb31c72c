"""Size-independent properties of the CUDA path at BASELINE.json's FULL sizes (config B: 4096 rays x 64+128; config E:
65,536 rays), where the CPU oracle would take minutes: sortedness, known answers of the reference's edge cases
(SURVEY.md 8c), linearity, chunk invariance, idempotence of the stash variant, RNG draw order."""
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def ops():
    from nerficg_b200 import ops
    return ops


def _gen(seed):
    return torch.Generator(device=DEV).manual_seed(seed)


# ---- K1 ------------------------------------------------------------------------------------------------------
def test_stratified_full_size_properties(ops):
    n, nc = 65536, 64
    det = ops.sample_stratified(n, nc, 2.0, 6.0, None, torch.device(DEV))
    assert (det[0].cpu() - torch.linspace(2.0, 6.0, nc)).abs().max() <= 5e-7    # torch.linspace to the last bit or one ulp (CPU FMA)
    assert torch.equal(det, det[:1].expand_as(det))                              # every ray gets the same depths
    u = torch.rand(n, nc, generator=_gen(1), device=DEV)
    z = ops.sample_stratified(n, nc, 2.0, 6.0, u, torch.device(DEV))
    mids = 0.5 * (det[:, 1:] + det[:, :-1])
    lo = torch.cat((det[:, :1], mids), -1)
    hi = torch.cat((mids, det[:, -1:]), -1)
    assert bool(((z >= lo) & (z <= hi)).all())                                  # one sample per stratum
    assert bool((z[:, 1:] >= z[:, :-1]).all())                                  # hence ascending
    z0 = ops.sample_stratified(n, nc, 2.0, 6.0, torch.zeros_like(u), torch.device(DEV))
    assert torch.equal(z0, lo)                                                   # u = 0 -> lower stratum borders exactly


# ---- K2 ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('nc,nf', [(64, 128), (128, 384)])
def test_importance_full_size_properties(ops, nc, nf):
    n = 65536
    g = _gen(2)
    z_c = ops.sample_stratified(n, nc, 2.0, 6.0, torch.rand(n, nc, generator=g, device=DEV), torch.device(DEV))
    w = torch.rand(n, nc, generator=g, device=DEV) * (torch.rand(n, nc, generator=g, device=DEV) < 0.3)
    u = torch.rand(n, nf, generator=g, device=DEV)
    merged, fine = ops.sample_importance(z_c, w, nf, u, True)
    assert merged.shape == (n, nc + nf)
    assert bool((merged[:, 1:] >= merged[:, :-1]).all())                         # exactly sorted (reference Renderer.py:70)
    ref_sorted = torch.sort(torch.cat((z_c, fine), -1), -1).values
    assert torch.equal(merged, ref_sorted)                                       # merge-by-rank == sort(cat()), bit for bit
    bins = 0.5 * (z_c[:, 1:] + z_c[:, :-1])
    assert bool(((fine >= bins[:, :1] - 1e-6) & (fine <= bins[:, -1:] + 1e-6)).all())  # fine samples stay inside the bin range
    # known answer: zero weights -> uniform pdf -> z = bins[0] + u * (bins[-1] - bins[0]) up to rounding
    z_det = ops.sample_stratified(n, nc, 2.0, 6.0, None, torch.device(DEV))
    _, fine0 = ops.sample_importance(z_det, torch.zeros_like(w), nf, u, True)
    b = 0.5 * (z_det[:, 1:] + z_det[:, :-1])
    expect = b[:, :1] + u * (b[:, -1:] - b[:, :1])
    assert (fine0 - expect).abs().max() <= 2e-5
    # deterministic mode: u = linspace(0, 1, nf) (reference utils.py:93) equals passing that noise explicitly
    # (torch.linspace's two-sided formula written out as separate multiply and subtract, so no FMA contraction differs)
    k = torch.arange(nf, device=DEV, dtype=torch.float32)
    step = torch.tensor(1.0, device=DEV) / float(nf - 1)
    lin = torch.where(k < nf // 2, step * k, 1.0 - step * (nf - 1 - k))
    m_det = ops.sample_importance(z_c, w, nf, None)
    m_lin = ops.sample_importance(z_c, w, nf, lin.expand(n, nf).contiguous())
    assert torch.equal(m_det, m_lin)


# ---- K5 / K6 ---------------------------------------------------------------------------------------------------
def _composite_inputs(n, s, seed):
    g = _gen(seed)
    z = torch.sort(2 + 4 * torch.rand(n, s, generator=g, device=DEV), -1).values
    sigma = torch.empty(n, s, device=DEV).exponential_(1.0, generator=g) * (torch.rand(n, s, generator=g, device=DEV) < 0.5)
    rgb = torch.rand(n, s, 3, generator=g, device=DEV)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, device=DEV), dim=-1) * 1.05
    return z, sigma, rgb, dirs


def test_composite_full_size_known_answers(ops):
    n, s = 65536, 192
    z, sigma, rgb, dirs = _composite_inputs(n, s, 3)
    bg = torch.tensor([1.0, 0.5, 0.25], device=DEV)
    pack = lambda sg, c: torch.cat((c, sg[..., None]), -1).contiguous()
    # zero density -> rgb = background, depth = 0, alpha = 0
    o_rgb, o_depth, o_alpha, o_w = ops.composite_forward(z, pack(torch.zeros_like(sigma), rgb), dirs, bg, True)
    assert torch.equal(o_rgb, bg.expand(n, 3)) and not o_depth.any() and not o_alpha.any() and not o_w.any()
    # any density on the last sample (delta = 1e10) -> alpha = 1 exactly
    sg = sigma.clone()
    sg[:, -1] = 0.5
    o_rgb, o_depth, o_alpha, o_w = ops.composite_forward(z, pack(sg, rgb), dirs, bg, True)
    assert torch.equal(o_alpha, torch.ones_like(o_alpha))
    assert (o_w.sum(-1, keepdim=True) - o_alpha).abs().max() <= 1e-5              # weights sum to alpha
    assert bool(((o_depth >= z[:, :1] - 1e-4) & (o_depth <= z[:, -1:] + 1e-4)).all())  # depth is a convex combination
    # colours are composited linearly: C(c1 + c2) - C(c2) = sum_i w_i c1_i (no background term left)
    r1 = ops.composite_forward(z, pack(sg, rgb), dirs, bg)[0]
    r0 = ops.composite_forward(z, pack(sg, torch.zeros_like(rgb)), dirs, bg)[0]
    assert ((r1 - r0) - (o_w[..., None] * rgb).sum(1)).abs().max() <= 2e-5
    # a subset of the rays evaluated alone gives the same bits (rays are independent: the sharding invariant of config C)
    part = ops.composite_forward(z[1000:3000].contiguous(), pack(sg, rgb)[1000:3000].contiguous(), dirs[1000:3000].contiguous(), bg)[0]
    assert torch.equal(part, r1[1000:3000])


def test_composite_backward_full_size_linearity_and_oracle_sample(ops):
    n, s = 65536, 192
    z, sigma, rgb, dirs = _composite_inputs(n, s, 4)
    bg = torch.ones(3, device=DEV)
    rs = torch.cat((rgb, sigma[..., None]), -1).contiguous()
    g = _gen(5)
    g1, g2 = torch.randn(n, 3, generator=g, device=DEV), torch.randn(n, 3, generator=g, device=DEV)
    d1 = ops.composite_backward(z, rs, dirs, bg, g1)
    d2 = ops.composite_backward(z, rs, dirs, bg, g2)
    d12 = ops.composite_backward(z, rs, dirs, bg, g1 + g2)
    # the backward is linear in the upstream gradient (the last sample's d(sigma) carries delta = 1e10 and only ever meets
    # ReLU'(0) = 0 downstream: it is left out)
    lin_err = (d1 + d2 - d12).abs() / d12.abs().clamp_min(1.0)
    assert lin_err[..., :3].max() <= 1e-4 and lin_err[:, :-1, 3].max() <= 1e-4
    # colour gradients are w_i * dL/drgb exactly
    w = ops.composite_forward(z, rs, dirs, bg, True)[3]
    assert (d1[..., :3] - w[..., None] * g1[:, None, :]).abs().max() <= 1e-6
    # a slice of the full-size launch against autograd of the CPU oracle
    sl = slice(4096, 4096 + 64)
    zz, sg, cc, dd = (t[sl].cpu() for t in (z, sigma, rgb, dirs))
    sg.requires_grad_(True)
    cc.requires_grad_(True)
    o_rgb, _, _, _ = O.composite(zz, dd, sg, cc, bg.cpu())
    (o_rgb * g1[sl].cpu()).sum().backward()
    assert (d1[sl, :, :3].cpu() - cc.grad).abs().max() <= 1e-5
    inner = d1[sl, :-1, 3].cpu() - sg.grad[:, :-1]
    assert (inner.abs() / sg.grad[:, :-1].abs().clamp_min(1.0)).max() <= 1e-4


# ---- K3 at config B ----------------------------------------------------------------------------------------------
def test_mlp_forward_full_size_invariants(ops):
    from nerficg_b200 import params
    n, s = 4096, 192
    g = torch.Generator().manual_seed(6)
    flat = (torch.rand(params.layout()[2], generator=g) - 0.5).mul(0.12).to(DEV)
    packed = ops.mlp_pack(flat)
    o = (torch.randn(n, 3, generator=g) * 0.1).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(DEV)
    z = torch.sort(2 + 4 * torch.rand(n, s, generator=g), -1).values.to(DEV)
    out = ops.mlp_forward(packed, flat, o, d, d, z)
    assert bool(torch.isfinite(out).all())
    assert bool(((out[..., :3] > 0) & (out[..., :3] < 1)).all()) and bool((out[..., 3] >= 0).all())   # sigmoid / ReLU ranges
    # idempotence: the training variant (activation stash) computes the same bits, run after run
    stash = torch.empty(ops.mlp_stash_bytes(n * s), dtype=torch.uint8, device=DEV)
    assert torch.equal(ops.mlp_forward(packed, flat, o, d, d, z, None, stash), out)
    assert torch.equal(ops.mlp_forward(packed, flat, o, d, d, z), out)
    # samples are independent: any ray range evaluated alone (other tile alignment, other CTA pair) gives the same bits
    a, b = 1000, 1000 + 1537
    part = ops.mlp_forward(packed, flat, o[a:b].contiguous(), d[a:b].contiguous(), d[a:b].contiguous(), z[a:b].contiguous())
    assert torch.equal(part, out[a:b])
    # density noise enters before the ReLU (reference Model.py:74-77)
    noise = torch.randn(n * s, 1, generator=g).to(DEV)
    noisy = ops.mlp_forward(packed, flat, o, d, d, z, noise)
    assert torch.equal(noisy[..., :3], out[..., :3])
    big = torch.full((n * s, 1), 1e4, device=DEV)
    raw = ops.mlp_forward(packed, flat, o, d, d, z, big)[..., 3] - 1e4                 # recovers the pre-activation
    assert (torch.relu(raw + noise.view(n, s)) - noisy[..., 3]).abs().max() <= 2e-2   # 1e4 + x rounds x to 1e-3


# ---- renderer: chunking and RNG draw order -------------------------------------------------------------------------
def test_renderer_chunking_and_rng_order():
    from nerficg_b200 import Framework
    Framework.setup(None, {'RENDERER.N_SAMPLES': 192, 'RENDERER.COARSE_RATIO': 0.3333333, 'RENDERER.RAY_BATCH_SIZE': 300,
                           'GLOBAL.LOG_LEVEL': 0})
    from nerficg_b200.Cameras import PerspectiveCamera, SharedCameraSettings
    from nerficg_b200.Datasets import RayBatch
    from nerficg_b200.Implementations import Methods
    model = Methods.get_model('NeRF', name='t')
    model.load_state_dict(O.init_state_dict(0), strict=True)
    renderer = Methods.get_renderer('NeRF', model)
    cam = PerspectiveCamera(shared_settings=SharedCameraSettings(torch.ones(3), 2.0, 6.0), width=100, height=100,
                            focal_x=138.889, focal_y=138.889)
    g = torch.Generator().manual_seed(7)
    n = 700
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    rays = RayBatch(origin=(-4.0 * d + 0.05 * torch.randn(n, 3, generator=g)).to(DEV), direction=(d * 1.05).to(DEV), view_direction=d.to(DEV))
    dev = torch.device(DEV)
    std = 0.5
    with torch.no_grad():
        torch.manual_seed(11)
        torch.cuda.manual_seed(11)
        out = renderer.render_rays(rays, cam, randomize_samples=True, random_noise_density=std)
        # the same draws in the reference's order: per chunk rand(n,Nc), randn(n*Nc,1), rand(n,Nf), randn(n*S,1)
        torch.manual_seed(11)
        torch.cuda.manual_seed(11)
        draws = []
        for m in (300, 300, 100):
            u_c = torch.rand((m, 64), dtype=torch.float32, device=dev)
            n_c = std * torch.randn((m * 64, 1), dtype=torch.float32, device=dev)
            u_f = torch.rand(m, 128, device=dev)
            n_f = std * torch.randn((m * 192, 1), dtype=torch.float32, device=dev)
            draws.append({'u_c': u_c, 'n_c': n_c, 'u_f': u_f, 'n_f': n_f})
        fed = renderer.render_rays(rays, cam, randomize_samples=True, random_noise_density=std, noise=draws)
        for k in out:
            assert out[k].shape[0] == n and torch.equal(out[k], fed[k]), k
        # chunking is invisible: one chunk of 700 with the concatenated draws gives the same bits
        renderer.RAY_BATCH_SIZE = 4096
        whole = renderer.render_rays(rays, cam, randomize_samples=True, random_noise_density=std, noise=[{
            'u_c': torch.cat([x['u_c'] for x in draws]), 'u_f': torch.cat([x['u_f'] for x in draws]),
            'n_c': torch.cat([x['n_c'].view(-1, 64) for x in draws]).reshape(-1, 1),
            'n_f': torch.cat([x['n_f'].view(-1, 192) for x in draws]).reshape(-1, 1)}])
        for k in out:
            assert torch.equal(whole[k], out[k]), k
