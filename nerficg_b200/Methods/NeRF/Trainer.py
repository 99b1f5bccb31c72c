"""NeRF trainer (reference src/Methods/NeRF/Trainer.py): Adam(lr=1) x LambdaLR(log-linear 5e-4 -> 5e-5), one
ray batch per iteration, MSE on fine + coarse colour.

Two equivalent ways to run an iteration:
  * ``training_iteration`` -- the reference's sequence (render_rays -> NeRFLoss -> backward -> step) through
    autograd; this is the drop-in path the base-class loop calls.
  * ``fused_step`` -- the same arithmetic with every kernel called directly on static buffers and the whole
    iteration captured in ONE CUDA graph (no per-op host overhead, SURVEY hard part 6); used by bench.py and by
    ``training_iteration`` when TRAINING.FUSED_STEP is on.  Under torch.distributed the flat gradient buffers are
    all-reduced by NCCL between backward and the optimiser step (data parallel).
"""
from __future__ import annotations

import os

import torch

from ... import Framework, dist, ops, params
from ...Logging import Logger
from ...profiling import stage
from ...Optim.FlatAdam import FlatAdam
from ...Optim.lr_utils import LRDecayPolicy
from ...Optim.Samplers import DatasetSampler, RandomImageSampler, RayPoolSampler
from ..Base.Trainer import BaseTrainer
from ..Base.utils import pre_training_callback, training_callback
from .Loss import NeRFLoss
from .Renderer import default_grad_scale


@Framework.Configurable.configure(
    NUM_ITERATIONS=500000,
    BATCH_SIZE=1024,
    SAMPLE_SINGLE_IMAGE=True,
    DENSITY_RANDOM_NOISE_STD=0.0,
    LR_INIT=5e-04,
    LR_FINAL=5e-05,
    LAMBDA_COLOR_LOSS=1.0,
    LAMBDA_ALPHA_LOSS=0.0,
    FUSED_STEP=True,
    FLAT_ADAM=True,
)
class NeRFTrainer(BaseTrainer):
    def __init__(self, **kwargs) -> None:
        super().__init__(**kwargs)
        device = next(self.model.parameters()).device
        # lr lives in a device tensor so that a captured optimiser step follows the schedule
        if device.type == 'cuda' and self.FLAT_ADAM:
            self.optimizer = FlatAdam(self.model.blocks(), lr=1.0)          # K7: one launch per flat block buffer
        elif device.type == 'cuda':
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=torch.tensor(1.0, device=device), fused=True, capturable=True)
        else:
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=1.0)
        self.lr_scheduler = _LambdaLR(
            self.optimizer, lr_lambda=LRDecayPolicy(lr_init=self.LR_INIT, lr_final=self.LR_FINAL, max_steps=self.NUM_ITERATIONS),
            last_epoch=self.model.num_iterations_trained - 1)
        self.loss = NeRFLoss(self.LAMBDA_COLOR_LOSS, self.LAMBDA_ALPHA_LOSS, self.model.coarse_nerf is not None)
        self.sampler_train = None
        self.sampler_val = None
        self._fused: dict[tuple, '_FusedStep'] = {}
        # data parallel: every rank starts from rank 0's weights whatever its own seed drew
        if dist.world_size() > 1:
            dist.broadcast_parameters_([b.flat_params for b in self.model.blocks()])

    @classmethod
    def _method_classes(cls):
        from .Model import NeRF
        from .Renderer import NeRFRenderer
        return NeRF, NeRFRenderer

    @pre_training_callback(priority=1000)
    @torch.no_grad()
    def init_samplers(self, _, dataset) -> None:
        make = (lambda ds: DatasetSampler(ds, random=True, img_sampler_cls=RandomImageSampler)) if self.SAMPLE_SINGLE_IMAGE \
            else (lambda ds: RayPoolSampler(ds, img_sampler_cls=RandomImageSampler))
        self.sampler_train = make(dataset.train())
        if self.RUN_VALIDATION:
            self.sampler_val = make(dataset.eval())

    @training_callback(priority=50)
    def training_iteration(self, _, dataset) -> None:
        self.model.train()
        self.loss.train()
        dataset.train()
        if self.sampler_train is None:   # resumed from a '.train' file: the pre-training callbacks do not run again (Base/Trainer.py:234-240)
            self.init_samplers(_, dataset)
            dataset.train()
        ray_batch = self.sampler_train.get(dataset=dataset, ray_batch_size=self.BATCH_SIZE)['ray_batch']
        camera = dataset.default_camera
        if self.FUSED_STEP and len(ray_batch) <= self.renderer.RAY_BATCH_SIZE:
            self.fused_step(ray_batch, camera)
            return
        outputs = self.renderer.render_rays(ray_batch, camera, randomize_samples=True,
                                            random_noise_density=self.DENSITY_RANDOM_NOISE_STD)
        loss = self.loss(outputs, ray_batch, camera.background_color)
        loss.backward()
        if dist.world_size() > 1:   # the autograd path exchanges gradients too (one all-reduce per parameter tensor)
            dist.allreduce_mean_([p.grad for p in self.model.parameters() if p.grad is not None])
        self.optimizer.step()
        self.optimizer.zero_grad()
        self.lr_scheduler.step()

    @training_callback(active='RUN_VALIDATION', priority=100)
    @torch.no_grad()
    def validation_iteration(self, _, dataset) -> None:
        self.model.eval()
        self.loss.eval()
        dataset.eval()
        ray_batch = self.sampler_val.get(dataset=dataset, ray_batch_size=self.BATCH_SIZE)['ray_batch']
        outputs = self.renderer.render_rays(ray_batch, dataset.default_camera)
        self.loss(outputs, ray_batch, dataset.default_camera.background_color)

    # ---- fused iteration ---------------------------------------------------------------------
    def fused_step(self, ray_batch, camera, use_graph: bool = True) -> torch.Tensor:
        """One full training iteration on ``ray_batch`` (single chunk).  Returns the loss as a device scalar."""
        n = len(ray_batch)
        # everything a captured step freezes is part of the key; at most two steps (each owns a multi-GB stash and a graph) are kept
        bg_t = camera.background_color
        cached = getattr(self, '_camera_key', None)     # (a device-resident background colour must not cost a sync per step)
        if cached is None or cached[0] is not camera or cached[1] is not bg_t:
            cached = self._camera_key = (camera, bg_t, tuple(float(c) for c in bg_t.flatten().tolist()))
        key = (n, float(camera.near_plane), float(camera.far_plane), cached[2],
               float(self.LAMBDA_COLOR_LOSS), float(self.LAMBDA_ALPHA_LOSS), float(self.DENSITY_RANDOM_NOISE_STD), bool(use_graph))
        step = self._fused.get(key)
        if step is None:
            while len(self._fused) >= 2:
                self._fused.pop(next(iter(self._fused)))
            step = self._fused[key] = _FusedStep(self, n, camera, use_graph)
        loss = step.run(ray_batch)
        if self.loss.activate_logging:   # the fused step bypasses the loss module: keep its running averages fed
            self.loss.log_fused(float(loss))
        self.lr_scheduler.step()
        return loss


class _LambdaLR:
    """torch.optim.lr_scheduler.LambdaLR semantics (lr = base_lr * lr_lambda(last_epoch), stepped once on
    construction) for optimisers whose lr is a DEVICE tensor: the new value is written with an asynchronous fill,
    where torch's scheduler would read the tensor back to the host every step."""

    def __init__(self, optimizer: torch.optim.Optimizer, lr_lambda, last_epoch: int = -1) -> None:
        self.optimizer, self.lr_lambda, self.last_epoch = optimizer, lr_lambda, last_epoch
        self.base_lrs = [float(g['lr']) for g in optimizer.param_groups]
        self._last_lr = list(self.base_lrs)
        self.step()

    def step(self) -> None:
        self.last_epoch += 1
        self._last_lr = [base * self.lr_lambda(self.last_epoch) for base in self.base_lrs]
        for group, lr in zip(self.optimizer.param_groups, self._last_lr):
            if isinstance(group['lr'], torch.Tensor):
                group['lr'].fill_(lr)
            else:
                group['lr'] = lr

    def get_last_lr(self) -> list[float]:
        return self._last_lr


class _FusedStep:
    """Static-shape training iteration: buffers are allocated once, the kernel sequence is captured in a CUDA graph
    after two eager warm-up runs, and later calls only copy the ray batch into place and replay."""

    def __init__(self, trainer: NeRFTrainer, n_rays: int, camera, use_graph: bool) -> None:
        self.t = trainer
        model, renderer = trainer.model, trainer.renderer
        self.n = n_rays
        self.nc, self.nf = renderer.n_samples_coarse_nerf, renderer.n_samples_nerf
        dev = next(model.parameters()).device
        self.dev = dev
        self.near, self.far = float(camera.near_plane), float(camera.far_plane)
        self.bg = camera.background_color.to(device=dev, dtype=torch.float32).contiguous()
        self.noise_std = float(trainer.DENSITY_RANDOM_NOISE_STD)
        self.blocks = model.blocks()
        f32 = dict(dtype=torch.float32, device=dev)
        self.origin, self.direction, self.view_direction = (torch.zeros(n_rays, 3, **f32) for _ in range(3))
        self.rgb_gt = torch.zeros(n_rays, 3, **f32)
        self.alpha_gt = torch.ones(n_rays, 1, **f32)
        self.loss_out = torch.zeros((), **f32)
        s_tot = self.nc + self.nf
        self.stash = [torch.empty(ops.mlp_stash_bytes(n_rays * s), dtype=torch.uint8, device=dev)
                      for s in ([self.nc, s_tot] if self.nc > 0 else [s_tot])]
        self.ws = torch.empty(ops.mlp_backward_workspace_bytes(n_rays * s_tot), dtype=torch.uint8, device=dev)
        self.packed = [torch.empty(ops.mlp_packed_bytes(), dtype=torch.uint8, device=dev) for _ in self.blocks]
        # persistent flat gradient buffers (slices of ONE allocation, so that both networks can be exchanged in a single
        # collective); parameter .grad fields are views into them
        sizes = [b.flat_params.numel() for b in self.blocks]
        self.grads_all = torch.zeros(sum(sizes), **f32)
        self.grads = list(torch.split(self.grads_all, sizes))
        if isinstance(trainer.optimizer, FlatAdam):
            trainer.optimizer.bind_flat_grads(self.grads)
        self.scale = default_grad_scale(n_rays)
        self.world = dist.world_size()
        self.flat_adam = isinstance(trainer.optimizer, FlatAdam)
        if self.flat_adam:
            # K7 reads grad * 1/world (the NCCL exchange stays a plain SUM) and clears the buffers behind the read
            trainer.optimizer.grad_mult = 1.0 / self.world
            trainer.optimizer.zero_bound_grads = True
        # how the gradients are exchanged under torch.distributed: 'overlap' = the fine network's all-reduce
        # on a side stream behind its wgrad + the coarse one at the end; 'merged' = one all-reduce of both networks at the end;
        # 'none' = no exchange at all (DIAGNOSTIC: the ranks diverge; separates the exchange cost from the max-over-ranks skew)
        self.exchange = os.environ.get('NERF_B200_DDP_EXCHANGE', 'overlap')
        if self.exchange not in ('overlap', 'merged', 'none'):
            raise ValueError(f'NERF_B200_DDP_EXCHANGE={self.exchange!r}: expected overlap, merged or none')
        self.side_stream = torch.cuda.Stream(device=dev) if self.world > 1 and self.exchange == 'overlap' else None
        self.graph = None
        self.use_graph = use_graph
        self.calls = 0

    def _bind_grads(self) -> None:
        for block, grad in zip(self.blocks, self.grads):
            for p, view in zip(block.ordered_parameters(), params.views(grad).values()):
                p.grad = view

    def _body(self) -> None:
        with stage('training step (K1..K8)'):
            self._body_impl()

    def _body_impl(self) -> None:
        t, n, dev = self.t, self.n, self.dev
        flats = [b.flat_params for b in self.blocks]
        for flat, packed in zip(flats, self.packed):
            ops.mlp_pack(flat, packed, with_backward=True)
        fine = len(self.blocks) - 1
        std = self.noise_std
        if self.nc > 0:
            u_c = torch.rand((n, self.nc), dtype=torch.float32, device=dev)
            n_c = std * torch.randn((n * self.nc, 1), dtype=torch.float32, device=dev) if std > 0 else None
            z_c = ops.sample_stratified(n, self.nc, self.near, self.far, u_c, dev)
            rs_c = ops.mlp_forward(self.packed[0], flats[0], self.origin, self.direction, self.view_direction, z_c, n_c, self.stash[0])
            rgb_c, _, alpha_c, w_c = ops.composite_forward(z_c, rs_c, self.direction, self.bg, want_weights=True)
            u_f = torch.rand(n, self.nf, device=dev)
            z = ops.sample_importance(z_c, w_c, self.nf, u_f)
        else:
            u_f = torch.rand((n, self.nf), dtype=torch.float32, device=dev)
            z = ops.sample_stratified(n, self.nf, self.near, self.far, u_f, dev)
        n_f = std * torch.randn((z.numel(), 1), dtype=torch.float32, device=dev) if std > 0 else None
        rs_f = ops.mlp_forward(self.packed[fine], flats[fine], self.origin, self.direction, self.view_direction, z, n_f, self.stash[fine])
        rgb, _, alpha, _ = ops.composite_forward(z, rs_f, self.direction, self.bg)
        # K8: NeRFLoss (reference Loss.py:26-43) and its gradient for both passes in one launch
        lc, la = float(t.LAMBDA_COLOR_LOSS), float(t.LAMBDA_ALPHA_LOSS)
        coarse = self.nc > 0
        _, g_rgb, g_rgb_c, g_alpha, g_alpha_c = ops.loss_mse(
            rgb, rgb_c if coarse else None, alpha.reshape(-1), alpha_c.reshape(-1) if coarse else None, self.rgb_gt,
            self.alpha_gt.reshape(-1), self.bg, lc, la, loss_out=self.loss_out)
        if not self.flat_adam:      # (K7 leaves the flat gradient buffers zeroed behind its read)
            for g in self.grads:
                g.zero_()
        d_rs = ops.composite_backward(z, rs_f, self.direction, self.bg, g_rgb, None, g_alpha, True, self.scale)
        ops.mlp_backward(self.grads[fine], d_rs, rs_f, self.stash[fine], self.ws, self.packed[fine], flats[fine], n, z.shape[1], self.scale)
        if self.world > 1 and coarse and self.exchange == 'overlap':
            # the fine network's gradients are reduced over NVLink while the coarse backward runs
            main = torch.cuda.current_stream()
            self.side_stream.wait_stream(main)
            with torch.cuda.stream(self.side_stream):
                dist.allreduce_sum_([self.grads[fine]])
        if coarse:
            d_rs_c = ops.composite_backward(z_c, rs_c, self.direction, self.bg, g_rgb_c, None, g_alpha_c, True, self.scale)
            ops.mlp_backward(self.grads[0], d_rs_c, rs_c, self.stash[0], self.ws, self.packed[0], flats[0], n, self.nc, self.scale)
        if self.world > 1:
            if self.exchange == 'none':
                pass
            elif self.exchange == 'merged' or not coarse:
                dist.allreduce_sum_([self.grads_all])
            else:
                dist.allreduce_sum_([self.grads[0]])
                torch.cuda.current_stream().wait_stream(self.side_stream)
            if not self.flat_adam:
                for g in self.grads:
                    g.mul_(1.0 / self.world)
        t.optimizer.step()

    @torch.no_grad()
    def run(self, ray_batch) -> torch.Tensor:
        self.origin.copy_(ray_batch.origin, non_blocking=True)
        self.direction.copy_(ray_batch.direction, non_blocking=True)
        self.view_direction.copy_(ray_batch.view_direction, non_blocking=True)
        self.rgb_gt.copy_(ray_batch.rgb, non_blocking=True)
        if ray_batch.alpha is not None:
            self.alpha_gt.copy_(ray_batch.alpha, non_blocking=True)
        self._bind_grads()
        self.calls += 1
        if not self.use_graph or self.calls <= 2:
            self._body()          # eager warm-up (also initialises Adam state and NCCL)
            return self.loss_out
        if self.graph is None:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    self._body()
                self.graph = graph
            except Exception as e:  # capture unsupported (e.g. a collective that cannot be captured): stay eager
                Logger.log_warning(f'CUDA graph capture of the training step failed, running eagerly: {e}')
                self.use_graph = False
                torch.cuda.synchronize()
                self._body()
                return self.loss_out
        self.graph.replay()
        return self.loss_out
