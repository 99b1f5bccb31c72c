"""Weighted loss container (reference src/Optim/Losses/{Base,utils}.py): named loss terms with weights and
running averages; a term with weight 0 contributes a constant 0 and is not evaluated (utils.py:54-57)."""
from __future__ import annotations

from typing import Any, Callable

import torch

from .. import Framework


class _Item:
    def __init__(self, name: str, fn: Callable, weight: float | None = None, is_loss: bool = True) -> None:
        self.name, self.fn, self.is_loss = name, fn, is_loss
        self.weight = max(0.0, weight) if weight is not None else 0.0
        self.reset()

    def reset(self) -> None:
        self._sum = [0.0, 0.0]
        self._n = [0, 0]

    def get_average(self):
        return [self._sum[i] / self._n[i] if self._n[i] else 0.0 for i in range(2)]

    def apply(self, train: bool, kwargs: dict[str, Any], accumulate: bool) -> torch.Tensor:
        if self.is_loss and self.weight <= 0.0:
            return torch.tensor(0.0)
        val = self.fn(**kwargs)
        if self.is_loss:
            val = val * self.weight
        if accumulate:
            i = 0 if train else 1
            self._sum[i] += float(val.detach())
            self._n[i] += 1
        return val


class BaseLoss(torch.nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.loss_metrics: list[_Item] = []
        self.quality_metrics: list[_Item] = []
        self.activate_logging: bool = bool(Framework.config.TRAINING.WANDB.ACTIVATE)

    def add_loss_metric(self, name: str, metric: Callable, weight: float = None) -> None:
        self.loss_metrics.append(_Item(name, metric, weight, True))

    def add_quality_metric(self, name: str, metric: Callable) -> None:
        self.quality_metrics.append(_Item(name, metric, None, False))

    def reset(self) -> None:
        for item in self.loss_metrics + self.quality_metrics:
            item.reset()

    def log_fused(self, total: float) -> None:
        """Running average of the training loss when the step ran as one captured graph (K8 returns only the weighted
        total, so it is booked on the first loss term; the per-term split needs the autograd path, TRAINING.FUSED_STEP=False)."""
        if self.loss_metrics:
            item = self.loss_metrics[0]
            item._sum[0] += float(total)
            item._n[0] += 1

    def forward(self, configurations: dict[str, dict[str, Any]]) -> torch.Tensor:
        try:
            if self.activate_logging:
                with torch.no_grad():
                    for m in self.quality_metrics:
                        m.apply(self.training, configurations[m.name], True)
            total = 0.0
            for m in self.loss_metrics:
                total = total + m.apply(self.training, configurations[m.name], self.activate_logging)
            return total
        except KeyError as e:
            raise Framework.LossError(f'missing argument configuration for loss {e}')
