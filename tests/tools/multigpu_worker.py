"""Worker of tests/test_multigpu.py (one process per GPU, launched by torchrun; NCCL over NVLink).

Checks (SURVEY.md section 4 "distributed tests", 8e):
  render   N-rank ray-sharded / view-sharded rendering == 1-rank rendering, bit for bit (no collective on the data path);
  ddp      one data-parallel step on N ranks x B rays == one single-GPU step on the concatenated N*B rays;
  fused    the captured training step with the overlapped all-reduce keeps every rank's weights identical.
Rank 0 prints one line `MULTIGPU_RESULT {json}`.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as td

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def main() -> None:
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    from nerficg_b200 import Framework, dist
    Framework.setup(None, {'RENDERER.N_SAMPLES': 192, 'RENDERER.COARSE_RATIO': 0.3333333, 'RENDERER.RAY_BATCH_SIZE': 4096,
                           'TRAINING.NUM_ITERATIONS': 1000, 'GLOBAL.LOG_LEVEL': 0}, device_index=local)
    dev = Framework.config.GLOBAL.DEFAULT_DEVICE
    import datetime
    td.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=90))
    from oracle import nerf_oracle as O
    from nerficg_b200.Datasets import RayBatch
    from nerficg_b200.Datasets.Synthetic import SyntheticLegoDataset
    from nerficg_b200.Implementations import Methods
    from nerficg_b200.Methods.NeRF import TRAINING_INSTANCE
    res: dict = {'world': world}

    sd0 = O.init_state_dict(3)
    ds = SyntheticLegoDataset(96, 96, 4, 4, seed=0, device=dev)       # identical on every rank
    ds.precompute_rays(['train'])

    def fresh():
        model = Methods.get_model('NeRF', name='t')
        model.load_state_dict(sd0, strict=True)
        return model, Methods.get_renderer('NeRF', model)

    # ---------------- render: shards == whole, bit for bit ----------------
    model, renderer = fresh()
    model.eval()
    views = ds.test()
    keys = ('rgb', 'depth', 'alpha', 'rgb_coarse', 'depth_coarse', 'alpha_coarse')
    with torch.no_grad():
        # (a) by contiguous ray ranges inside one view
        rays = views[0].get_rays()
        mine = dist.shard_range(len(rays), rank, world)
        part = renderer.render_rays(rays[mine.start:mine.stop], views[0].camera)
        # (b) by view
        my_views = dist.shard_range(len(views), rank, world)
        imgs = {i: renderer.render_image(views[i]) for i in my_views}
        full = renderer.render_rays(rays, views[0].camera) if rank == 0 else None
        full_imgs = {i: renderer.render_image(views[i]) for i in range(len(views))} if rank == 0 else None
    dataset_train = ds.train()
    equal_rays, equal_views = True, True
    for k in keys:
        sizes = [len(dist.shard_range(len(rays), r, world)) for r in range(world)]
        bufs = [torch.empty((s, part[k].shape[1]), device=dev) for s in sizes]
        td.all_gather(bufs, part[k].contiguous())                       # test plumbing only: the data path has no collective
        if rank == 0:
            equal_rays &= torch.equal(torch.cat(bufs), full[k])
    for i in range(len(views)):
        owner = next(r for r in range(world) if i in dist.shard_range(len(views), r, world))
        t = imgs[i]['rgb'].contiguous() if owner == rank else torch.empty((96, 96, 3), device=dev)
        td.broadcast(t, src=owner)
        if rank == 0:
            equal_views &= torch.equal(t, full_imgs[i]['rgb'])
    res['render_ray_shards_bit_equal'], res['render_view_shards_bit_equal'] = bool(equal_rays), bool(equal_views)

    # ---------------- ddp: N ranks x B rays == 1 rank x N*B rays ----------------
    b_per = 512
    g = torch.Generator().manual_seed(5)
    pool = ds.ray_collection['train'].all_rays
    ids = torch.randint(0, len(pool), (world * b_per,), generator=g).to(dev)
    u_c, u_f = torch.rand(world * b_per, 64, generator=g).to(dev), torch.rand(world * b_per, 128, generator=g).to(dev)
    cam = ds.default_camera
    bg = cam.background_color.to(dev)

    from nerficg_b200.Methods.NeRF.Loss import NeRFLoss
    loss_fn = NeRFLoss(1.0, 0.0, True)

    def grads_of(model, renderer, sl):
        """render_rays -> NeRFLoss -> backward on the rays ids[sl]; NO collective in here (rank 0 also calls it alone)."""
        batch = pool[ids[sl]]
        out = renderer.render_rays(batch, cam, randomize_samples=True, noise=[{'u_c': u_c[sl].contiguous(), 'u_f': u_f[sl].contiguous()}])
        loss = loss_fn(out, batch, bg)
        loss.backward()
        return None, loss

    model, renderer = fresh()
    trainer, loss = grads_of(model, renderer, slice(rank * b_per, (rank + 1) * b_per))
    dist.allreduce_mean_([p.grad for p in model.parameters()])          # what NeRFTrainer.training_iteration does under DDP
    flat_ddp = torch.cat([p.grad.flatten() for p in model.parameters()])
    if rank == 0:
        model1, renderer1 = fresh()
        _, loss1 = grads_of(model1, renderer1, slice(0, world * b_per))
        flat_one = torch.cat([p.grad.flatten() for p in model1.parameters()])
        res['ddp_grad_rel_l2_vs_single_gpu'] = float((flat_ddp - flat_one).norm() / flat_one.norm())
        res['ddp_grad_max_abs'] = float((flat_ddp - flat_one).abs().max())
    lsum = loss.detach().clone()
    td.all_reduce(lsum)
    if rank == 0:
        res['ddp_loss_mean'], res['single_loss'] = float(lsum / world), float(loss1)

    # ---------------- fused: captured step with the overlapped all-reduce ----------------
    torch.manual_seed(100 + rank)                                        # different weights per rank before the broadcast
    model = Methods.get_model('NeRF', name='t')
    renderer = Methods.get_renderer('NeRF', model)
    trainer = TRAINING_INSTANCE(model=model, renderer=renderer)          # broadcasts rank 0's weights
    w0 = torch.cat([b.flat_params for b in model.blocks()]).clone()
    first = [torch.empty_like(w0) for _ in range(world)]
    td.all_gather(first, w0)
    for it in range(5):                                                   # 2 eager + capture + 2 replays
        batch = pool[torch.randint(0, len(pool), (1024,), device=dev)]
        trainer.fused_step(batch, cam)
    torch.cuda.synchronize()
    w = torch.cat([b.flat_params for b in model.blocks()])
    allw = [torch.empty_like(w) for _ in range(world)]
    td.all_gather(allw, w)
    if rank == 0:
        res['broadcast_equal_at_start'] = all(torch.equal(first[0], x) for x in first)
        res['fused_weights_identical_across_ranks'] = all(torch.equal(allw[0], x) for x in allw)
        res['fused_weights_moved'] = float((w - w0).abs().max()) > 0
        res['fused_graph_captured'] = next(iter(trainer._fused.values())).graph is not None
        res['finite'] = bool(torch.isfinite(w).all())
        print('MULTIGPU_RESULT ' + json.dumps(res), flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)    # (no NCCL teardown: collectives live inside a captured graph, see bench.py _leave)


if __name__ == '__main__':
    main()
