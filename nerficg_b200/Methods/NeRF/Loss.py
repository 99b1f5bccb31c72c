"""NeRF loss (reference src/Methods/NeRF/Loss.py:26-43): MSE on fine + coarse colour (and alpha when
LAMBDA_ALPHA_LOSS > 0) against the ground truth composited over the background; PSNR as quality metric.
This torch module serves the autograd path; the captured training step uses K8 (csrc/loss.cu) for the same arithmetic."""
from __future__ import annotations

import torch

from ...Datasets.utils import RayBatch, apply_background_color
from ...Optim.Losses import BaseLoss


def peak_signal_noise_ratio(preds: torch.Tensor, target: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    return 10.0 * torch.log10(data_range ** 2 / torch.mean((preds - target) ** 2))


class NeRFLoss(BaseLoss):
    """One (colour MSE, alpha MSE, PSNR) triple per rendering pass; metric names carry the pass suffix."""
    _PASSES = (('', ''), ('_coarse', '_Coarse'))   # (suffix of the renderer's output keys, suffix of the metric names)

    def __init__(self, lambda_color: float, lambda_alpha: float, requires_coarse_losses: bool) -> None:
        super().__init__()
        self.coarse_losses = bool(requires_coarse_losses)
        self._active = self._PASSES if self.coarse_losses else self._PASSES[:1]
        for _, tag in self._active:
            self.add_loss_metric(f'L2_Color{tag}', torch.nn.functional.mse_loss, lambda_color)
            self.add_loss_metric(f'L2_Alpha{tag}', torch.nn.functional.mse_loss, lambda_alpha)
            self.add_quality_metric(f'PSNR{tag}', peak_signal_noise_ratio)

    def forward(self, outputs: dict[str, torch.Tensor], rays: RayBatch, bg_color: torch.Tensor) -> torch.Tensor:
        alpha_target = rays.alpha if rays.alpha is not None else torch.ones_like(outputs['alpha'], requires_grad=False)
        color_target = apply_background_color(rays.rgb, alpha_target, bg_color, is_chw=False)
        terms = {}
        for key, tag in self._active:
            rgb, alpha = outputs[f'rgb{key}'], outputs[f'alpha{key}']
            terms[f'L2_Color{tag}'] = dict(input=rgb, target=color_target)
            terms[f'L2_Alpha{tag}'] = dict(input=alpha, target=alpha_target)
            terms[f'PSNR{tag}'] = dict(preds=rgb, target=color_target, data_range=1.0)
        return super().forward(terms)
