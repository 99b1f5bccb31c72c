"""K7 (flat-buffer Adam) against torch.optim.Adam, the reference's optimiser (src/Methods/NeRF/Trainer.py:32-37)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def fw():
    from nerficg_b200 import Framework
    Framework.setup(None, {'RENDERER.N_SAMPLES': 192, 'RENDERER.COARSE_RATIO': 0.3333333, 'GLOBAL.LOG_LEVEL': 0})
    return Framework


def _model():
    from nerficg_b200.Implementations import Methods
    return Methods.get_model('NeRF', name='t')


def test_flat_adam_matches_torch_adam(fw):
    from nerficg_b200.Optim.FlatAdam import FlatAdam
    torch.manual_seed(3)
    a = _model()
    b = _model()
    b.load_state_dict(a.state_dict())
    opt_a = FlatAdam(a.blocks(), lr=1.0)
    opt_b = torch.optim.Adam(b.parameters(), lr=1.0)
    g = torch.Generator(device=DEV).manual_seed(0)
    for it in range(6):
        lr = 5e-4 * (0.9 ** it)
        opt_a.param_groups[0]['lr'].fill_(lr)
        opt_b.param_groups[0]['lr'] = lr
        for pa, pb in zip(a.parameters(), b.parameters()):
            grad = torch.randn(pa.shape, generator=g, device=DEV) * (10.0 ** -(it % 4))
            if it == 2:
                grad = grad * (torch.rand(pa.shape, generator=g, device=DEV) < 0.5)   # exact zeros: eps path
            pa.grad, pb.grad = grad.clone(), grad.clone()
        opt_a.step()
        opt_b.step()
        opt_a.zero_grad()
        opt_b.zero_grad()
    for (name, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=2e-5, atol=1e-7), (name, (pa - pb).abs().max())
    # padding floats of the flat buffers never move
    for blk in a.blocks():
        flat = blk.flat_params.clone()
        for v in __import__('nerficg_b200').params.views(flat).values():
            v.zero_()
        assert flat.abs().max() == 0
    # checkpoint layout = torch.optim.Adam's
    sa, sb = opt_a.state_dict(), opt_b.state_dict()
    assert len(sa['state']) == len(sb['state']) == 48
    for i in range(48):
        assert float(sa['state'][i]['step']) == float(sb['state'][i]['step']) == 6
        assert torch.allclose(sa['state'][i]['exp_avg'], sb['state'][i]['exp_avg'], rtol=1e-4, atol=1e-7)
        assert torch.allclose(sa['state'][i]['exp_avg_sq'], sb['state'][i]['exp_avg_sq'], rtol=1e-4, atol=1e-10)
    # round trip into a fresh optimiser, one more identical step
    c = _model()
    c.load_state_dict(a.state_dict())
    opt_c = FlatAdam(c.blocks(), lr=1.0)
    opt_c.load_state_dict(copy.deepcopy(sa))
    for pa, pc in zip(a.parameters(), c.parameters()):
        grad = torch.randn(pa.shape, generator=g, device=DEV)
        pa.grad, pc.grad = grad.clone(), grad.clone()
    opt_a.step()
    opt_c.step()
    for pa, pc in zip(a.parameters(), c.parameters()):
        assert torch.equal(pa, pc)
