"""Stage-wise parity of the CUDA kernels (through the C-ABI) against the oracle and the reference goldens."""
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def ops():
    from nerficg_b200 import ops
    return ops


@pytest.mark.parametrize('mode,n,k', [(0, 256, 256), (0, 128, 64), (0, 64, 128), (1, 256, 128), (1, 64, 64), (1, 128, 256),
                                      (2, 256, 256)])
def test_umma_selftest(ops, mode, n, k):
    g = torch.Generator().manual_seed(mode * 100 + n + k)
    a = torch.randn(128, k, generator=g)
    b = torch.randn(n, k, generator=g)
    q = (lambda t: t.bfloat16().double()) if mode == 2 else (lambda t: t.half().double())
    ref = q(a) @ q(b).T
    out = ops.selftest_umma(a.to(DEV), b.to(DEV), mode).cpu()
    assert torch.allclose(out.double(), ref, rtol=1e-4, atol=1e-3), (out.double() - ref).abs().max()


@pytest.mark.parametrize('n,k', [(256, 256), (128, 64), (64, 128), (256, 64)])
def test_umma_cta_pair_selftest(ops, n, k):
    g = torch.Generator().manual_seed(n + k)
    a = torch.randn(256, k, generator=g)
    b = torch.randn(n, k, generator=g)
    ref = a.half().double() @ b.half().double().T
    out = ops.selftest_umma2(a.to(DEV), b.to(DEV)).cpu()
    assert torch.allclose(out.double(), ref, rtol=1e-4, atol=1e-3), (out.double() - ref).abs().max()


# K2 vs the CPU oracle.  SURVEY 8c allows a 1e-5 fraction of outliers "at the denom < 1e-5 branch", but that branch is not the
# only place where the reference's arithmetic is ill-conditioned: a sample that falls into a near-empty bin (pdf ~ 1e-5 / sum)
# inherits the cdf's rounding error times 1/pdf * bin width, i.e. ONE ulp of the cdf (6e-8) moves it by ~4e-4, and torch's own
# CPU and CUDA cumsum/sum differ by ulps.  What IS exact is the inverse-CDF property |F(z) - u| <= 2e-5 under the oracle's own
# cdf (asserted for EVERY sample below).  The bound here is the measured outlier fraction (gpurun_out/parity_measured.jsonl,
# DESIGN.md section 2) with a 2x margin; test inputs (weights = rand^8) are chosen to stress near-empty bins.
K2_OUTLIER_FRACTION = 5e-3    # measured on B200 (round 2): 2.7e-4 .. 2.6e-3 over these cases, max |dz| 5e-2 at the `denom < 1e-5` branch


def _record_k2(name, frac, worst, numel):
    import json, os
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/parity_measured.jsonl', 'a') as f:
        f.write(json.dumps({'test': 'K2_' + name, 'outlier_fraction_gt_2e-5': frac, 'max_abs': worst, 'samples': numel}) + '\n')


@pytest.mark.parametrize('n,k', [(256, 128), (256, 64), (128, 128), (256, 256)])
def test_umma_cta_pair_mn_major_selftest(ops, n, k):
    """cta_group::2 MMAs with BOTH operands MN-major (stored reduction-major): the descriptor flavour of the fused
    weight-gradient path (dW += dY^T X over the samples)."""
    g = torch.Generator().manual_seed(3 * n + k)
    a = torch.randn(256, k, generator=g)
    b = torch.randn(n, k, generator=g)
    ref = a.half().double() @ b.half().double().T
    out = ops.selftest_umma2(a.to(DEV), b.to(DEV), mn_major=True).cpu()
    assert torch.allclose(out.double(), ref, rtol=1e-4, atol=1e-3), (out.double() - ref).abs().max()


def test_stratified_golden(ops, golden):
    g = golden('stratified')
    z = ops.sample_stratified(g['n'], g['nc'], g['near'], g['far'], g['u'].to(DEV), torch.device(DEV)).cpu()
    assert (z - g['z_rand']).abs().max() <= 1e-6
    z = ops.sample_stratified(g['n'], g['nc'], g['near'], g['far'], None, torch.device(DEV)).cpu()
    assert (z - g['z_det']).abs().max() <= 1e-6


@pytest.mark.parametrize('n,nc', [(1, 64), (1000, 64), (257, 128), (33, 7), (5, 1)])
def test_stratified_oracle(ops, n, nc):
    u = torch.rand(n, nc, generator=torch.Generator().manual_seed(n))
    z = ops.sample_stratified(n, nc, 2.0, 6.0, u.to(DEV), torch.device(DEV)).cpu()
    assert (z - O.stratified_depths(n, nc, 2.0, 6.0, u)).abs().max() <= 1e-6


def test_importance_golden(ops, golden):
    g = golden('importance')
    merged, fine = ops.sample_importance(g['z_coarse'].to(DEV), g['w_coarse'].to(DEV), g['nf'], g['u'].to(DEV), True)
    # 1e-5-class agreement, except where the reference algorithm itself is discontinuous (SURVEY hard part 7):
    # an empty bin of an opaque ray has cdf mass 1e-5/(1+62e-5), one fp32 ulp away from the `denom < 1e-5`
    # threshold, so whether such a bin interpolates or collapses to its left edge depends on the last bit of the
    # cdf (torch CPU vs torch CUDA differ there too).  Affected samples move by at most one bin (< 0.1).
    def check_close(a, b):
        err = (a - b).abs()
        frac, worst = (err > 2e-5).float().mean().item(), err.max().item()
        _record_k2('importance_golden', frac, worst, err.numel())
        assert frac <= K2_OUTLIER_FRACTION and worst <= 0.1, (frac, worst)
    check_close(fine.cpu(), g['zf_rand'])
    check_close(merged.cpu(), g['merged_rand'])
    merged, fine = ops.sample_importance(g['z_coarse'].to(DEV), g['w_coarse'].to(DEV), g['nf'], None, True)
    check_close(fine.cpu(), g['zf_det'])


@pytest.mark.parametrize('n,nc,nf', [(512, 64, 128), (100, 64, 192), (64, 128, 384), (7, 16, 9), (3, 4, 5)])
def test_importance_oracle(ops, n, nc, nf):
    g = torch.Generator().manual_seed(nc * 7 + nf)
    zc = O.stratified_depths(n, nc, 2.0, 6.0, torch.rand(n, nc, generator=g))
    w = torch.rand(n, nc, generator=g) ** 8
    w[0] = 0
    u = torch.rand(n, nf, generator=g)
    merged, fine = ops.sample_importance(zc.to(DEV), w.to(DEV), nf, u.to(DEV), True)
    ref = O.importance_depths(zc, w, nf, u)
    err = (fine.cpu() - ref).abs()
    # samples in near-empty bins amplify cdf rounding by 1/pdf and the `denom < 1e-5` branch is a
    # discontinuity (SURVEY hard part 7): positions agree to 2e-5 except for a small fraction of outliers,
    # and EVERY sample satisfies the inverse-CDF property |F(z) - u| <= 2e-5 under the oracle's own CDF
    _record_k2(f'importance_oracle[{n},{nc},{nf}]', (err > 2e-5).float().mean().item(), err.max().item(), err.numel())
    assert (err > 2e-5).float().mean() <= K2_OUTLIER_FRACTION, err.max()
    edges = 0.5 * (zc[:, :-1] + zc[:, 1:]).double()
    v = w[:, 1:-1].double() + 1e-5
    cdf = torch.cat((torch.zeros(n, 1, dtype=torch.double), torch.cumsum(v / v.sum(-1, keepdim=True), -1)), -1)
    zf = fine.cpu().double()
    j = (torch.searchsorted(edges, zf.contiguous(), right=True) - 1).clamp(0, nc - 3)
    e0, e1 = torch.gather(edges, 1, j), torch.gather(edges, 1, j + 1)
    c0, c1 = torch.gather(cdf, 1, j), torch.gather(cdf, 1, j + 1)
    f_of_z = c0 + (zf - e0) / (e1 - e0) * (c1 - c0)
    assert (f_of_z - u.double()).abs().max() <= 2e-5
    m = merged.cpu()
    assert torch.all(m[:, 1:] >= m[:, :-1])                       # sortedness
    assert torch.equal(torch.sort(torch.cat((zc, fine.cpu()), -1), -1).values, m)  # same multiset as its own fine samples


def _composite_case(n, s, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(2 + 4 * torch.rand(n, s, generator=g), -1).values
    sigma = (torch.rand(n, s, generator=g) < 0.4).float() * (-torch.log(torch.rand(n, s, generator=g))) * 8
    if n > 1:
        sigma[0] = 0
        sigma[1, -1] = 2.0
    color = torch.rand(n, s, 3, generator=g)
    d = torch.randn(n, 3, generator=g)
    return z, sigma, color, d


def test_composite_golden(ops, golden):
    g = golden('composite')
    rs = torch.cat((g['color'], g['sigma'][..., None]), -1).to(DEV)
    rgb, depth, alpha, w = ops.composite_forward(g['z'].to(DEV), rs, g['dirs'].to(DEV), g['bg'].to(DEV), True)
    for got, ref in ((rgb, g['rgb']), (depth, g['depth']), (alpha, g['alpha']), (w, g['w'])):
        assert (got.cpu() - ref).abs().max() <= 1e-5
    assert alpha[1].item() == 1.0 and alpha[0].item() == 0.0 and depth[0].item() == 0.0
    d = ops.composite_backward(g['z'].to(DEV), rs, g['dirs'].to(DEV), g['bg'].to(DEV), g['g_rgb'].to(DEV), None,
                               g['g_alpha'].reshape(-1).to(DEV)).cpu()
    assert torch.allclose(d[..., :3], g['d_color'], rtol=1e-4, atol=1e-6)
    assert torch.allclose(d[..., 3], g['d_sigma'], rtol=1e-4, atol=1e-6 * g['d_sigma'].abs().max().item())


@pytest.mark.parametrize('n,s', [(1, 1), (3, 31), (64, 64), (500, 192), (129, 256), (40, 512), (9, 100)])
def test_composite_oracle(ops, n, s):
    z, sigma, color, d = _composite_case(n, s, n * 1000 + s)
    bg = torch.tensor([1.0, 0.5, 0.25])
    sr, cr = sigma.clone().requires_grad_(True), color.clone().requires_grad_(True)
    rgb_o, depth_o, alpha_o, w_o = O.composite(z, d, sr, cr, bg)
    rs = torch.cat((color, sigma[..., None]), -1).to(DEV)
    rgb, depth, alpha, w = ops.composite_forward(z.to(DEV), rs, d.to(DEV), bg.to(DEV), True)
    assert (rgb.cpu() - rgb_o).abs().max() <= 1e-5
    assert (alpha.cpu() - alpha_o).abs().max() <= 1e-5
    assert (w.cpu() - w_o).abs().max() <= 1e-5
    assert torch.allclose(depth.cpu(), depth_o.detach(), rtol=1e-4, atol=1e-5)
    g = torch.Generator().manual_seed(s)
    g_rgb, g_a, g_d = torch.randn(n, 3, generator=g), torch.randn(n, 1, generator=g), torch.randn(n, 1, generator=g)
    # the reference's depth has no usable gradient on (nearly) empty rays (0/0 under torch.where,
    # see the comment at src/Methods/NeRF/utils.py:129): only feed depth gradients to solid rays
    g_d = g_d * (alpha_o.detach() > 1e-2)
    ((rgb_o * g_rgb).sum() + (alpha_o * g_a).sum() + (depth_o * g_d).sum()).backward()
    dd = ops.composite_backward(z.to(DEV), rs, d.to(DEV), bg.to(DEV), g_rgb.to(DEV), g_d.reshape(-1).to(DEV),
                                g_a.reshape(-1).to(DEV)).cpu()
    assert torch.allclose(dd[..., :3], cr.grad, rtol=1e-4, atol=1e-6)
    assert torch.isfinite(dd).all()
    # the reference's autograd is NaN on empty rays once depth is in the graph (0/0): compare finite rows only
    ok = torch.isfinite(sr.grad).all(dim=-1)
    assert ok.float().mean() > 0.5
    # the last sample carries the 1e10 interval: check it separately so it cannot mask the others
    for sl in (slice(0, s - 1), slice(s - 1, s)):
        if sl.stop > sl.start:
            ref = sr.grad[ok][:, sl]
            assert (dd[ok][:, sl, 3] - ref).abs().max() <= 2e-4 * (ref.abs().max().item() + 1e-12)
    # relu_mask zeroes the density gradient where sigma <= 0 and scales everything
    dm = ops.composite_backward(z.to(DEV), rs, d.to(DEV), bg.to(DEV), g_rgb.to(DEV), None, None, True, 4.0).cpu()
    assert torch.all(dm[..., 3][sigma <= 0] == 0)


def test_composite_no_background_and_empty(ops):
    z, sigma, color, d = _composite_case(6, 64, 5)
    rs = torch.cat((color, sigma[..., None]), -1).to(DEV)
    rgb, depth, alpha, _ = ops.composite_forward(z.to(DEV), rs, d.to(DEV), None)
    rgb_o, depth_o, alpha_o, _ = O.composite(z, d, sigma, color, None)
    assert (rgb.cpu() - rgb_o).abs().max() <= 1e-5
    e = ops.composite_forward(z[:0].to(DEV), rs[:0], d[:0].to(DEV), None)
    assert e[0].shape == (0, 3)


@pytest.mark.parametrize('n,lam_a,coarse', [(4096, 0.0, True), (1000, 0.3, True), (77, 0.5, False), (1, 0.0, True), (70000, 0.1, True)])
def test_loss_kernel_oracle(ops, n, lam_a, coarse):
    """K8 against the oracle's NeRFLoss (reference Loss.py:26-43) and autograd of it."""
    g = torch.Generator().manual_seed(n)
    mk = lambda *shape: torch.rand(*shape, generator=g)
    out = {'rgb': mk(n, 3).requires_grad_(True), 'alpha': mk(n, 1).requires_grad_(True)}
    if coarse:
        out |= {'rgb_coarse': mk(n, 3).requires_grad_(True), 'alpha_coarse': mk(n, 1).requires_grad_(True)}
    rgb_gt, alpha_gt, bg = mk(n, 3) * 1.2 - 0.1, (mk(n, 1) * 1.4 - 0.2).clamp(0, 1), torch.tensor([1.0, 0.6, 0.2])
    ref = O.nerf_loss(out, rgb_gt, alpha_gt, bg, lambda_color=0.7, lambda_alpha=lam_a)
    ref.backward()
    dev = lambda t: None if t is None else t.detach().to(DEV)
    loss, g_rgb, g_rgb_c, g_alpha, g_alpha_c = ops.loss_mse(dev(out['rgb']), dev(out.get('rgb_coarse')), dev(out['alpha']).reshape(-1),
                                                           dev(out['alpha_coarse']).reshape(-1) if coarse else None, dev(rgb_gt),
                                                           dev(alpha_gt).reshape(-1), dev(bg), 0.7, lam_a)
    assert abs(loss.item() - ref.item()) <= 1e-6 + 1e-5 * abs(ref.item())
    assert (g_rgb.cpu() - out['rgb'].grad).abs().max() <= 1e-7 + 1e-5 * out['rgb'].grad.abs().max()
    if coarse:
        assert (g_rgb_c.cpu() - out['rgb_coarse'].grad).abs().max() <= 1e-7 + 1e-5 * out['rgb_coarse'].grad.abs().max()
    if lam_a > 0:
        assert (g_alpha.cpu() - out['alpha'].grad.reshape(-1)).abs().max() <= 1e-7 + 1e-5 * out['alpha'].grad.abs().max()
        if coarse:
            assert (g_alpha_c.cpu() - out['alpha_coarse'].grad.reshape(-1)).abs().max() <= 1e-7
    else:
        assert g_alpha is None and g_alpha_c is None


def test_generate_rays_golden(ops, golden):
    """K0 against View.get_rays of the reference (tests/golden/rays.pt) and against the oracle on a third camera."""
    for case in golden('rays'):
        args = (case['c2w'].numpy(), case['width'], case['height'], case['focal_x'], case['focal_y'], case['center_x'], case['center_y'])
        o, d, v = ops.generate_rays(*args, case['pixel_ids'].to(DEV), torch.device(DEV))
        assert torch.equal(o.cpu(), case['origin'])
        assert (d.cpu() - case['direction']).abs().max() <= 1e-6
        assert (v.cpu() - case['view_direction']).abs().max() <= 1e-6
        # all pixels (no id list) == the same rays in row-major order
        o_all, d_all, v_all = ops.generate_rays(*args, None, torch.device(DEV))
        assert d_all.shape == (case['width'] * case['height'], 3)
        assert torch.equal(d_all[case['pixel_ids'].to(DEV)], d) and torch.equal(v_all[case['pixel_ids'].to(DEV)], v)
    c2w = golden('rays')[0]['c2w']
    ref = O.camera_rays(c2w, 800, 800, 1111.11, 1111.11, 400.0, 400.0)
    got = ops.generate_rays(c2w.numpy(), 800, 800, 1111.11, 1111.11, 400.0, 400.0, None, torch.device(DEV))   # config B/C image size
    for r, g_ in zip(ref, got):
        assert (g_.cpu() - r).abs().max() <= 1e-6


def test_gather_rays(ops):
    g = torch.Generator().manual_seed(3)
    n_pool, n = 100000, 4096
    src = {k: torch.rand(n_pool, 3, generator=g).to(DEV) for k in ('origin', 'direction', 'view_direction', 'rgb')}
    src['alpha'] = torch.rand(n_pool, 1, generator=g).to(DEV)
    ids = torch.randint(0, n_pool, (n,), generator=g).to(DEV)
    dst = {k: torch.empty((n,) + tuple(t.shape[1:]), device=DEV) for k, t in src.items()}
    ops.gather_rays(dst, src, ids)
    for k in src:
        assert torch.equal(dst[k], src[k][ids]), k
    partial = {'origin': torch.empty(n, 3, device=DEV)}
    ops.gather_rays(partial, src, ids)                       # any subset of the fields
    assert torch.equal(partial['origin'], src['origin'][ids])
