import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nerficg_b200 import Framework
from oracle import nerf_oracle as O
DEV = 'cuda:0'
Framework.setup(None, {'RENDERER.N_SAMPLES': 192, 'RENDERER.COARSE_RATIO': 0.3333333, 'RENDERER.RAY_BATCH_SIZE': 8192, 'GLOBAL.LOG_LEVEL': 0})
Framework.config.TRAINING.NUM_ITERATIONS = 1000
from nerficg_b200.Implementations import Methods
from nerficg_b200.Datasets.Synthetic import SyntheticLegoDataset
from nerficg_b200.Methods.NeRF import TRAINING_INSTANCE

# (a) RNG: eager vs graph
def draw():
    a = torch.rand((512, 64), dtype=torch.float32, device=DEV)
    b = torch.rand(512, 128, device=DEV)
    return a, b
torch.manual_seed(7); ea, eb = draw()
torch.manual_seed(7); _ = draw()
g = torch.cuda.CUDAGraph()
torch.manual_seed(7)
with torch.cuda.graph(g):
    ga, gb = draw()
torch.manual_seed(7)
g.replay(); torch.cuda.synchronize()
print('rng first equal', torch.equal(ea, ga), 'second equal', torch.equal(eb, gb))

ds = SyntheticLegoDataset(64, 64, 2, 1, device=DEV)
ds.precompute_rays(['train'])
batch = ds.ray_collection['train'][0][torch.arange(0, 4096, 8, device=DEV)]
res = []
for fused in (False, True):
    model = Methods.get_model('NeRF', name='t'); model.load_state_dict(O.init_state_dict(1))
    renderer = Methods.get_renderer('NeRF', model)
    trainer = TRAINING_INSTANCE(model=model, renderer=renderer)
    torch.manual_seed(100)
    if fused:
        step = trainer._fused[len(batch)] = __import__('nerficg_b200.Methods.NeRF.Trainer', fromlist=['_FusedStep'])._FusedStep(trainer, len(batch), ds.default_camera, False)
        # run body without optimizer effect: snapshot grads after
        step.run(batch)
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    else:
        out = renderer.render_rays(batch, ds.default_camera, randomize_samples=True)
        trainer.loss(out, batch, ds.default_camera.background_color).backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    res.append(grads)
for k in res[0]:
    a, b = res[0][k], res[1][k]
    rel = ((a - b).norm() / (a.norm() + 1e-30)).item()
    if rel > 1e-4:
        print(k, rel, a.norm().item(), b.norm().item())
print('done')
