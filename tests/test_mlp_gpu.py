"""K3/K4 parity: fused tcgen05 MLP against the oracle (fp32) with teacher-forced inputs."""
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
def net():
    from nerficg_b200 import ops, params
    sd = O.init_state_dict(0)
    flat = {p: params.flatten_state_dict(sd, p, DEV) for p in ('nerf.', 'coarse_nerf.')}
    packed = {p: ops.mlp_pack(f) for p, f in flat.items()}
    return sd, flat, packed


def test_points_golden(ops, net, golden):
    """NeRFBlock.forward on raw points (S = 1, origin = point, dir = 0): vs reference outputs."""
    sd, flat, packed = net
    g = golden('mlp')
    n = g['pts'].shape[0]
    zero = torch.zeros(n, 3, device=DEV)
    for prefix, ks, kc in (('nerf.', 'sigma', 'rgb'), ('coarse_nerf.', 'sigma_coarse', 'rgb_coarse')):
        out = ops.mlp_forward(packed[prefix], flat[prefix], g['pts'].to(DEV), zero, g['dirs'].to(DEV),
                              torch.zeros(n, 1, device=DEV)).cpu().reshape(n, 4)
        assert (out[:, :3] - g[kc]).abs().max() <= 1e-3
        assert (out[:, 3:] - g[ks]).abs().max() <= 1e-3 + 2e-3 * g[ks].abs().max()


@pytest.mark.parametrize('n_rays,s', [(1, 1), (3, 64), (100, 192), (37, 77), (1500, 64), (700, 192)])
def test_forward_vs_oracle(ops, net, n_rays, s):
    sd, flat, packed = net
    g = torch.Generator().manual_seed(n_rays + s)
    o = torch.randn(n_rays, 3, generator=g) * 2
    d = torch.randn(n_rays, 3, generator=g)
    vd = torch.nn.functional.normalize(d, dim=-1)
    z = torch.sort(2 + 4 * torch.rand(n_rays, s, generator=g), -1).values
    noise = torch.randn(n_rays * s, generator=g)
    x = o[:, None] + d[:, None] * z[..., None]
    sig, rgb = O.mlp_forward(sd, 'nerf.', x.reshape(-1, 3), vd[:, None].expand_as(x).reshape(-1, 3), noise[:, None])
    out = ops.mlp_forward(packed['nerf.'], flat['nerf.'], o.to(DEV), d.to(DEV), vd.to(DEV), z.to(DEV), noise.to(DEV))
    out = out.cpu().reshape(-1, 4)
    assert (out[:, :3] - rgb).abs().max() <= 1e-3, (out[:, :3] - rgb).abs().max()
    assert (out[:, 3:] - sig).abs().max() <= 1e-3 + 2e-3 * sig.abs().max(), (out[:, 3:] - sig).abs().max()
    # training variant (stash) must produce identical outputs
    stash = torch.empty(ops.mlp_stash_bytes(n_rays * s), dtype=torch.uint8, device=DEV)
    out_t = ops.mlp_forward(packed['nerf.'], flat['nerf.'], o.to(DEV), d.to(DEV), vd.to(DEV), z.to(DEV), noise.to(DEV), stash)
    assert torch.equal(out_t.cpu().reshape(-1, 4), out)
