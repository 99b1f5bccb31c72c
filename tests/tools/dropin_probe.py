"""Drop-in probe: runs the UNMODIFIED reference (/root/reference, CPU mode) with this package injected as its 'NeRF' method and
prints one JSON line of observations.  TEST INFRASTRUCTURE (build container only -- the reference tree does not travel to the
GPU box); tests/test_reference_dropin.py runs it in a subprocess because importing the reference puts its top-level modules
(Framework, Logging, Implementations, Methods, ...) into sys.modules.

Call sequence exercised = what the reference's scripts/train.py does (src/Implementations.py:43-65, src/Framework.py:73-108):
    Framework.setup(...)  ->  Implementations.Methods.get_training_instance(METHOD_TYPE)  ->  Datasets.get_dataset  ->  samplers
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))


def main() -> None:
    from blender_scene import write_scene
    from oracle.ref_loader import load_reference

    obs: dict = {}
    ref = load_reference(n_samples=192, coarse_ratio=0.3333333)            # reference Framework.load_config + CPU mode
    RF = ref['Framework']
    ref_method = ref['method']                                              # the reference's own NeRF plugin triple
    # a host run with non-default settings (what `-c cfg.yaml KEY=VAL ...` produces)
    RF.config.TRAINING.BATCH_SIZE = 2048
    RF.config.TRAINING.LR_INIT = 1.0e-3
    RF.config.TRAINING.NUM_ITERATIONS = 1234
    RF.config.TRAINING.MODEL_NAME = 'probe'
    RF.config.RENDERER.RAY_BATCH_SIZE = 4096
    import Implementations as RI                                            # the reference's registry

    import nerficg_b200.Framework as OF
    import nerficg_b200.Implementations as B200
    B200.install_into_reference(RI)                                         # the documented injection
    obs['config_is_host'] = OF.config is RF.config
    obs['directories_is_host'] = OF.Directories is RF.Directories

    model = RI.Methods.get_model('NeRF', name='probe')
    renderer = RI.Methods.get_renderer('NeRF', model)
    trainer = RI.Methods.get_training_instance('NeRF')
    obs['classes'] = [type(model).__module__, type(renderer).__module__, type(trainer).__module__]
    obs['renderer'] = dict(N_SAMPLES=renderer.N_SAMPLES, COARSE_RATIO=renderer.COARSE_RATIO, RAY_BATCH_SIZE=renderer.RAY_BATCH_SIZE,
                           n_coarse=renderer.n_samples_coarse_nerf, n_fine=renderer.n_samples_nerf)
    obs['trainer'] = dict(BATCH_SIZE=trainer.BATCH_SIZE, LR_INIT=trainer.LR_INIT, NUM_ITERATIONS=trainer.NUM_ITERATIONS,
                          MODEL_NAME=trainer.model.model_name, lr0=trainer.lr_scheduler.get_last_lr()[0],
                          keys=sorted(k for k in trainer.__dict__ if k.isupper()))
    obs['model_device'] = str(next(model.parameters()).device)
    obs['default_device'] = str(RF.config.GLOBAL.DEFAULT_DEVICE)
    obs['model_keys'] = dict(HIERARCHICAL=model.HIERARCHICAL, N_LAYERS=model.N_LAYERS, INPUT_SKIPS=list(model.INPUT_SKIPS))
    # the host rebinding its config (a second Framework.load_config) must be followed
    old_cfg = RF.config
    RF.load_config(RF.Directories.CONFIG_DIR / 'nerf_lego.yaml', True, {'RENDERER.N_SAMPLES': '96', 'RENDERER.COARSE_RATIO': '0.5'})
    RF.config.GLOBAL.GPU_INDICES = None
    RF.config.GLOBAL.DEFAULT_DEVICE = torch.device('cpu')
    r2 = RI.Methods.get_renderer('NeRF', model)
    obs['rebound'] = dict(N_SAMPLES=r2.N_SAMPLES, n_coarse=r2.n_samples_coarse_nerf, followed=OF.config is RF.config and OF.config is not old_cfg)

    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        # ---- reference dataset / View / RayBatch objects flowing into our trainer (Trainer.py:51-63) ----
        write_scene(tmp / 'scene')
        RF.config.DATASET.PATH = str(tmp / 'scene')
        RF.config.DATASET.NORMALIZE_CUBE = None
        dataset = RI.Datasets.get_dataset('NeRF', str(tmp / 'scene'))
        obs['dataset_class'] = type(dataset).__module__
        trainer.init_samplers(0, dataset)
        dataset.train()
        got = trainer.sampler_train.get(dataset=dataset, ray_batch_size=16)
        rb = got['ray_batch']
        obs['ray_batch'] = dict(cls=type(rb).__module__ + '.' + type(rb).__name__, n=len(rb), origin=list(rb.origin.shape),
                                view_direction=list(rb.view_direction.shape), rgb=list(rb.rgb.shape), alpha=list(rb.alpha.shape))
        obs['view_class'] = type(got['view']).__module__
        cam = dataset.default_camera
        obs['camera'] = dict(near=float(cam.near_plane), far=float(cam.far_plane), bg=[float(c) for c in cam.background_color])
        # everything in front of the first kernel launch accepts the reference's objects; the launch itself needs a B200
        try:
            renderer.render_rays(rb, cam, randomize_samples=True)
            obs['render_on_cpu'] = 'ran'
        except Exception as e:  # noqa: BLE001
            obs['render_on_cpu'] = type(e).__name__
        try:
            trainer.FUSED_STEP = False
            trainer.training_iteration(0, dataset)
            obs['train_on_cpu'] = 'ran'
        except Exception as e:  # noqa: BLE001
            obs['train_on_cpu'] = type(e).__name__

        # ---- checkpoints both ways (Base/Model.py:60-111) ----
        B200.uninstall_from_reference(RI)
        RI.Methods.modules['NeRF'] = ref_method
        torch.manual_seed(7)
        ref_model = ref_method.MODEL('ref_written').build()
        ref_model.num_iterations_trained = 321
        ref_model.save(tmp / 'ref.pt')
        B200.install_into_reference(RI)
        ours = RI.Methods.get_model('NeRF', checkpoint=str(tmp / 'ref.pt'))
        sd_r, sd_o = ref_model.state_dict(), ours.state_dict()
        obs['ckpt_ref_to_ours'] = dict(cls=type(ours).__module__, same_keys=sorted(sd_r) == sorted(sd_o),
                                       equal=all(torch.equal(sd_r[k], sd_o[k].cpu()) for k in sd_r),
                                       iters=ours.num_iterations_trained, name=ours.model_name)
        ours.save(tmp / 'ours.pt')
        back = ref_method.MODEL.load(str(tmp / 'ours.pt'))
        sd_b = back.state_dict()
        obs['ckpt_ours_to_ref'] = dict(cls=type(back).__module__, equal=all(torch.equal(sd_r[k], sd_b[k]) for k in sd_r),
                                       iters=back.num_iterations_trained)
        # ---- '.train' resume through the reference's accessor (Implementations.py:57-62) ----
        trainer.model.num_iterations_trained = 77
        trainer.save(tmp / 'state.train')
        resumed = RI.Methods.get_training_instance('NeRF', checkpoint=str(tmp / 'state.train'))
        obs['resume'] = dict(cls=type(resumed).__module__, iters=resumed.model.num_iterations_trained, BATCH_SIZE=resumed.BATCH_SIZE,
                             last_epoch=resumed.lr_scheduler.last_epoch,
                             equal=all(torch.equal(a, b) for a, b in zip(trainer.model.state_dict().values(), resumed.model.state_dict().values())))
    print('DROPIN_PROBE ' + json.dumps(obs))


if __name__ == '__main__':
    main()
