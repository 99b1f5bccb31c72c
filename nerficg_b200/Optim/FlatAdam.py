"""Adam on the flat parameter buffers of the NeRF blocks (C-ABI K7, ``nerf_adam_tick`` / ``nerf_adam_update``).

Same update rule, hyper-parameters and ``state_dict`` layout as ``torch.optim.Adam`` (the reference's optimiser,
src/Methods/NeRF/Trainer.py:32-37): per parameter ``step``, ``exp_avg``, ``exp_avg_sq`` -- here views into one flat
buffer per block, so a step is one 28 B/parameter stream per block instead of torch's 48-tensor multi-tensor launch
(measured 70 us per block).  The step counter, bias corrections and learning rate are device scalars: the whole step can
be captured in a CUDA graph and the schedule is written with an asynchronous fill.
"""
from __future__ import annotations

import torch

from .. import _lib, params as P


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, blocks, lr=1.0, betas=(0.9, 0.999), eps=1e-8) -> None:
        self.blocks = list(blocks)
        flats = [b.flat_params for b in self.blocks]
        device = flats[0].device
        _lib.require_device(flats[0])
        lr_t = lr if isinstance(lr, torch.Tensor) else torch.tensor(float(lr), dtype=torch.float32, device=device)
        super().__init__([p for b in self.blocks for p in b.ordered_parameters()], dict(lr=lr_t, betas=betas, eps=eps))
        self.exp_avg = [torch.zeros_like(f) for f in flats]
        self.exp_avg_sq = [torch.zeros_like(f) for f in flats]
        self.device_state = torch.zeros(4, dtype=torch.float32, device=device)   # {step, 1-b1^t, sqrt(1-b2^t), -}
        self._flat_grads: list[torch.Tensor | None] = [None] * len(self.blocks)
        self.grad_mult = 1.0          # folded into the gradient read (1/world_size under data parallelism)
        self.zero_bound_grads = False  # clear the bound flat gradient buffers behind the read (zero_grad folded into the step)

    def bind_flat_grads(self, grads: list[torch.Tensor]) -> None:
        """Registers the flat gradient buffers the fused training step accumulates into (parameter .grad fields are
        views of them); without it ``step`` gathers the per-parameter .grad tensors."""
        self._flat_grads = list(grads)

    def _gather_grad(self, i: int) -> torch.Tensor:
        block = self.blocks[i]
        flat = torch.zeros_like(block.flat_params)
        for p, view in zip(block.ordered_parameters(), P.views(flat).values()):
            if p.grad is not None:
                view.copy_(p.grad)
        return flat

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        group = self.param_groups[0]
        b1, b2 = group['betas']
        lr = group['lr']
        if not isinstance(lr, torch.Tensor):
            lr = group['lr'] = torch.tensor(float(lr), dtype=torch.float32, device=self.device_state.device)
        lib, stream = _lib.load(), _lib.stream_ptr()
        _lib.check(lib.nerf_adam_tick(self.device_state.data_ptr(), float(b1), float(b2), stream), 'nerf_adam_tick')
        for i, block in enumerate(self.blocks):
            flat = block.flat_params
            g = self._flat_grads[i]
            first = block.ordered_parameters()[0]
            bound = g is not None and first.grad is not None and first.grad.data_ptr() == g.data_ptr()
            if not bound:
                g = self._gather_grad(i)
            _lib.check(lib.nerf_adam_update_ex(flat.data_ptr(), self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr(), g.data_ptr(),
                                               lr.data_ptr(), self.device_state.data_ptr(), float(b1), float(b2), float(group['eps']),
                                               float(self.grad_mult), int(bound and self.zero_bound_grads), flat.numel(), stream),
                       'nerf_adam_update_ex')
        return loss

    # ---- torch.optim.Adam-compatible checkpoints ------------------------------------------------------------
    def state_dict(self) -> dict:
        state, idx = {}, 0
        step = self.device_state[0].detach().clone()
        for i, block in enumerate(self.blocks):
            m, v = P.views(self.exp_avg[i]), P.views(self.exp_avg_sq[i])
            for name in m:
                state[idx] = {'step': step.clone(), 'exp_avg': m[name].clone(), 'exp_avg_sq': v[name].clone()}
                idx += 1
        g = self.param_groups[0]
        return {'state': state, 'param_groups': [{'lr': float(g['lr']), 'betas': g['betas'], 'eps': g['eps'], 'weight_decay': 0,
                                                  'amsgrad': False, 'maximize': False, 'params': list(range(idx))}]}

    def load_state_dict(self, sd: dict) -> None:
        idx = 0
        for i, block in enumerate(self.blocks):
            m, v = P.views(self.exp_avg[i]), P.views(self.exp_avg_sq[i])
            for name in m:
                st = sd['state'].get(idx)
                if st is not None:
                    m[name].copy_(st['exp_avg'])
                    v[name].copy_(st['exp_avg_sq'])
                    self.device_state[0] = float(st['step'])
                idx += 1
        g = sd['param_groups'][0]
        self.param_groups[0]['lr'].fill_(float(g['lr']))
        self.param_groups[0]['betas'], self.param_groups[0]['eps'] = tuple(g['betas']), g['eps']
        # bias corrections are recomputed by the next tick from the restored step
