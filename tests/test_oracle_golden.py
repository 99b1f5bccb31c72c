"""Pins oracle/nerf_oracle.py to outputs of the unmodified reference (tests/golden, made by oracle/make_golden.py)."""
import math

import torch

from oracle import nerf_oracle as O


def test_stratified(golden):
    g = golden('stratified')
    z = O.stratified_depths(g['n'], g['nc'], g['near'], g['far'], g['u'])
    assert torch.equal(z, g['z_rand'])
    assert torch.equal(O.stratified_depths(g['n'], g['nc'], g['near'], g['far']), g['z_det'])


def test_composite_forward_backward(golden):
    g = golden('composite')
    s, c = g['sigma'].clone().requires_grad_(True), g['color'].clone().requires_grad_(True)
    rgb, depth, alpha, w = O.composite(g['z'], g['dirs'], s, c, g['bg'])
    for a, b in ((rgb, g['rgb']), (depth, g['depth']), (alpha, g['alpha']), (w, g['w'])):
        assert torch.equal(a.detach(), b)
    ((rgb * g['g_rgb']).sum() + (alpha * g['g_alpha']).sum()).backward()
    assert torch.allclose(s.grad, g['d_sigma'], rtol=1e-6, atol=0)
    assert torch.equal(c.grad, g['d_color'])
    # known answers: empty ray -> background, depth 0, alpha 0; opaque last sample -> alpha exactly 1
    assert torch.equal(rgb[0].detach(), g['bg']) and depth[0].item() == 0.0 and alpha[0].item() == 0.0
    assert alpha[1].item() == 1.0


def test_importance_and_merge(golden):
    g = golden('importance')
    zf = O.importance_depths(g['z_coarse'], g['w_coarse'], g['nf'], g['u'])
    assert torch.equal(zf, g['zf_rand'])
    assert torch.equal(O.importance_depths(g['z_coarse'], g['w_coarse'], g['nf']), g['zf_det'])
    assert torch.equal(O.merge_depths(g['z_coarse'], zf), g['merged_rand'])
    # zero weights -> every bin has equal mass: piecewise-linear map of u through the bin edges
    import numpy as np
    e = 0.5 * (g['z_coarse'][2, :-1] + g['z_coarse'][2, 1:])
    lin = np.interp(g['u'][2].double().numpy(), np.linspace(0, 1, e.numel()), e.double().numpy())
    assert np.allclose(zf[2].numpy(), lin, atol=2e-5)


def test_encoding_and_mlp(golden):
    g = golden('mlp')
    sd = O.init_state_dict(g['seed'])
    checksum = torch.stack([sd[k].double().abs().sum() for k in sorted(sd)]).sum()
    assert checksum.item() == g['checksum'].item(), 'torch CPU generator drifted: regenerate goldens'
    assert torch.equal(O.frequency_encoding(g['pts'], 10), g['enc_pos'])
    assert torch.equal(O.frequency_encoding(g['dirs'], 4), g['enc_dir'])
    x = torch.tensor([[0.1, 0.2, 0.3]])
    kat = O.frequency_encoding(x, 10)
    assert torch.equal(kat, g['kat_010203'])
    assert kat.shape == (1, 63) and abs(kat[0, 3].item() - math.cos(0.1)) < 1e-7 and abs(kat[0, 13].item() - math.sin(0.1)) < 1e-7
    for prefix, ks, kc in (('nerf.', 'sigma', 'rgb'), ('coarse_nerf.', 'sigma_coarse', 'rgb_coarse')):
        s, c = O.mlp_forward(sd, prefix, g['pts'], g['dirs'])
        assert torch.allclose(s, g[ks], rtol=1e-5, atol=1e-6)
        assert torch.allclose(c, g[kc], rtol=1e-5, atol=1e-6)


def _oracle_render(g, sd, randomize):
    outs = []
    n = g['o'].shape[0]
    for i, lo in enumerate(range(0, n, g['chunk'])):
        sl = slice(lo, min(lo + g['chunk'], n))
        dr = g['draws'][i] if randomize else {}
        outs.append(O.render_rays(sd, g['o'][sl], g['d'][sl], g['v'][sl], 2.0, 6.0, g['bg'], 64, 128,
                                  dr.get('u_c'), dr.get('u_f'), dr.get('n_c'), dr.get('n_f')))
    return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}


def test_render_loss_grads(golden):
    g = golden('render')
    sd = {k: v.clone().requires_grad_(v.dim() <= 2 and 'frequency' not in k) for k, v in O.init_state_dict(g['seed']).items()}
    out = _oracle_render(g, sd, True)
    for k, ref in g['out'].items():
        assert torch.allclose(out[k], ref, rtol=1e-5, atol=1e-6), k
    loss = O.nerf_loss(out, g['rgb_gt'], g['alpha_gt'], g['bg'])
    assert abs(loss.item() - g['loss'].item()) < 1e-6
    loss.backward()
    for k, ref in g['grad_norm'].items():
        assert abs(sd[k].grad.norm().item() - ref.item()) <= 1e-4 * ref.item() + 1e-9, k
        assert torch.allclose(sd[k].grad.flatten()[:48], g['grad_head'][k], rtol=1e-3, atol=1e-7), k
    with torch.no_grad():
        det = _oracle_render(g, sd, False)
    for k, ref in g['out_det'].items():
        assert torch.allclose(det[k], ref, rtol=1e-5, atol=1e-6), k


def test_lr_schedule(golden):
    g = golden('lr')
    for it, lr in zip(g['its'], g['lr']):
        assert abs(O.lr_factor(it, 5e-4, 5e-5, 500000) - lr) <= 1e-12


def test_camera_rays(golden):
    """Ray generation of the oracle against View.get_rays of the reference (oracle/make_golden_rays.py)."""
    for case in golden('rays'):
        o, d, v = O.camera_rays(case['c2w'], case['width'], case['height'], case['focal_x'], case['focal_y'], case['center_x'],
                                case['center_y'], case['pixel_ids'])
        assert torch.equal(o, case['origin'])
        assert (d - case['direction']).abs().max() <= 1e-6
        assert (v - case['view_direction']).abs().max() <= 1e-6
        assert (v.norm(dim=-1) - 1).abs().max() <= 1e-6


def test_oracle_matches_reference_in_the_trained_regime(golden):
    """The oracle replayed at weights the REFERENCE trained itself (its own checkpoint file, held-out PSNR > 20 dB, sigma_max ~ 40):
    deterministic render_rays, the randomised pass with the reference's noise, the loss and the leading gradient elements."""
    import torch
    from pathlib import Path
    from oracle import nerf_oracle as O
    g = golden('trained_render')
    ckpt = torch.load(Path(__file__).resolve().parent / 'golden' / 'ref_trained_checkpoint.pt', map_location='cpu', weights_only=False)
    sd = ckpt['model_state_dict']
    assert g['psnr_heldout'] > 20.0 and g['sigma_max'] > 20.0 and ckpt['num_iterations_trained'] == g['steps']
    out = O.render_rays(sd, g['o'], g['d'], g['v'], 2.0, 6.0, g['bg'], 64, 128)
    for k, ref in g['out_det'].items():
        assert (out[k] - ref).abs().max() <= 2e-5, k
    assert (out['z'] - g['z']).abs().max() <= 1e-5
    leaf = {k: v.clone().requires_grad_('frequency' not in k) for k, v in sd.items()}
    out = O.render_rays(leaf, g['o'], g['d'], g['v'], 2.0, 6.0, g['bg'], 64, 128, g['draws'][0]['u_c'], g['draws'][0]['u_f'])
    loss = O.nerf_loss(out, g['rgb_gt'], g['alpha_gt'], g['bg'])
    assert abs(float(loss) - float(g['loss'])) <= 1e-5 * float(g['loss'])
    loss.backward()
    for k, ref in g['grad_head'].items():
        got = leaf[k].grad.flatten()[:ref.numel()]
        assert (got - ref).norm() <= 1e-3 * max(float(ref.norm()), 0.05 * float(g['grad_norm'][k])) + 1e-12, k
