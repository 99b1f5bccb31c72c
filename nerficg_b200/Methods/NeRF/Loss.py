"""NeRF loss (reference src/Methods/NeRF/Loss.py:26-43): MSE on fine + coarse colour (and alpha when
LAMBDA_ALPHA_LOSS > 0) against the ground truth composited over the background; PSNR as quality metric."""
from __future__ import annotations

import torch

from ...Datasets.utils import RayBatch, apply_background_color
from ...Optim.Losses import BaseLoss


def peak_signal_noise_ratio(preds: torch.Tensor, target: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    return 10.0 * torch.log10(data_range ** 2 / torch.mean((preds - target) ** 2))


class NeRFLoss(BaseLoss):
    def __init__(self, lambda_color: float, lambda_alpha: float, requires_coarse_losses: bool) -> None:
        super().__init__()
        mse = torch.nn.functional.mse_loss
        self.coarse_losses = bool(requires_coarse_losses)
        self.add_loss_metric('L2_Color', mse, lambda_color)
        self.add_loss_metric('L2_Alpha', mse, lambda_alpha)
        self.add_quality_metric('PSNR', peak_signal_noise_ratio)
        if self.coarse_losses:
            self.add_loss_metric('L2_Color_Coarse', mse, lambda_color)
            self.add_loss_metric('L2_Alpha_Coarse', mse, lambda_alpha)
            self.add_quality_metric('PSNR_Coarse', peak_signal_noise_ratio)

    def forward(self, outputs: dict[str, torch.Tensor], rays: RayBatch, bg_color: torch.Tensor) -> torch.Tensor:
        alpha_gt = torch.ones_like(outputs['alpha'], requires_grad=False) if rays.alpha is None else rays.alpha
        color_gt = apply_background_color(rays.rgb, alpha_gt, bg_color, is_chw=False)
        losses = {'L2_Color': {'input': outputs['rgb'], 'target': color_gt},
                  'L2_Alpha': {'input': outputs['alpha'], 'target': alpha_gt},
                  'PSNR': {'preds': outputs['rgb'], 'target': color_gt, 'data_range': 1.0}}
        if self.coarse_losses:
            losses |= {'L2_Color_Coarse': {'input': outputs['rgb_coarse'], 'target': color_gt},
                       'L2_Alpha_Coarse': {'input': outputs['alpha_coarse'], 'target': alpha_gt},
                       'PSNR_Coarse': {'preds': outputs['rgb_coarse'], 'target': color_gt, 'data_range': 1.0}}
        return super().forward(losses)
