"""CPU: the documented injection into the UNMODIFIED reference -- config propagation, reference dataset /
View / RayBatch objects into our trainer, checkpoints both ways, '.train' resume.  Needs /root/reference (build container);
runs tests/tools/dropin_probe.py in a subprocess because importing the reference claims top-level module names."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path('/root/reference/src/Methods/NeRF/Renderer.py')


@pytest.fixture(scope='module')
def probe():
    if not REFERENCE.exists():
        pytest.skip('reference checkout not present (GPU box): the drop-in probe runs in the build container')
    r = subprocess.run([sys.executable, str(ROOT / 'tests' / 'tools' / 'dropin_probe.py')], capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith('DROPIN_PROBE ')]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(lines[-1][len('DROPIN_PROBE '):])


def test_host_config_reaches_the_plugin_classes(probe):
    """reference src/Framework.py:73-108: Configurable reads the HOST framework's config, including KEY=VAL overrides."""
    assert probe['config_is_host'] and probe['directories_is_host']
    assert probe['classes'] == ['nerficg_b200.Methods.NeRF.Model', 'nerficg_b200.Methods.NeRF.Renderer', 'nerficg_b200.Methods.NeRF.Trainer']
    r = probe['renderer']
    assert (r['N_SAMPLES'], r['RAY_BATCH_SIZE'], r['n_coarse'], r['n_fine']) == (192, 4096, 64, 128)
    assert abs(r['COARSE_RATIO'] - 0.3333333) < 1e-9
    t = probe['trainer']
    assert (t['BATCH_SIZE'], t['NUM_ITERATIONS'], t['MODEL_NAME']) == (2048, 1234, 'probe')
    assert t['LR_INIT'] == pytest.approx(1e-3) and t['lr0'] == pytest.approx(1e-3)      # LambdaLR starts at LR_INIT
    for key in ('NUM_ITERATIONS', 'BATCH_SIZE', 'SAMPLE_SINGLE_IMAGE', 'DENSITY_RANDOM_NOISE_STD', 'LR_INIT', 'LR_FINAL',
                'LAMBDA_COLOR_LOSS', 'LAMBDA_ALPHA_LOSS', 'LOAD_CHECKPOINT', 'MODEL_NAME', 'RUN_VALIDATION', 'BACKUP', 'WANDB'):
        assert key in t['keys'], key                                                     # Trainer.py:17-26, Base/Trainer.py:31-74
    assert probe['model_device'] == probe['default_device'] == 'cpu'                     # GLOBAL.DEFAULT_DEVICE of the host is honoured
    assert probe['model_keys'] == {'HIERARCHICAL': True, 'N_LAYERS': 8, 'INPUT_SKIPS': [5]}


def test_host_rebinding_its_config_is_followed(probe):
    """load_config REBINDS the module global (Framework.py:163-176): later constructions see the new object."""
    assert probe['rebound'] == {'N_SAMPLES': 96, 'n_coarse': 48, 'followed': True}


def test_reference_dataset_objects_flow_into_our_trainer(probe):
    """src/Methods/NeRF/Trainer.py:51-63 with the reference's own BaseDataset / View / RayBatch."""
    assert probe['dataset_class'] == 'Datasets.NeRF' and probe['view_class'] == 'Datasets.utils'
    rb = probe['ray_batch']
    assert rb['cls'] == 'Datasets.utils.RayBatch' and rb['n'] == 16
    assert rb['origin'] == [16, 3] and rb['view_direction'] == [16, 3] and rb['rgb'] == [16, 3] and rb['alpha'] == [16, 1]
    assert probe['camera'] == {'near': 2.0, 'far': 6.0, 'bg': [1.0, 1.0, 1.0]}
    # no CPU fallback: with the reference's objects everything runs up to the first kernel launch, which refuses the CPU
    assert probe['render_on_cpu'] == 'NativeLibraryError' and probe['train_on_cpu'] == 'NativeLibraryError'


def test_checkpoints_cross_both_ways_and_train_resume(probe):
    """src/Methods/Base/Model.py:60-111 (file written by the reference loads here and vice versa), Implementations.py:57-65."""
    a, b, c = probe['ckpt_ref_to_ours'], probe['ckpt_ours_to_ref'], probe['resume']
    assert a == {'cls': 'nerficg_b200.Methods.NeRF.Model', 'same_keys': True, 'equal': True, 'iters': 321, 'name': 'ref_written'}
    assert b == {'cls': 'Methods.NeRF.Model', 'equal': True, 'iters': 321}
    assert c == {'cls': 'nerficg_b200.Methods.NeRF.Trainer', 'iters': 77, 'BATCH_SIZE': 2048, 'last_epoch': 77, 'equal': True}
