"""Base renderer (reference src/Methods/Base/Renderer.py): model type check, ``render_image`` contract,
test-set rendering with 8-bit PSNR.  SSIM/LPIPS of the reference need torchmetrics + VGG weights (not part of
the hot path, not installable offline) and are out of scope; PSNR follows the same 8-bit round trip."""
from __future__ import annotations

from abc import ABC, abstractmethod
from pathlib import Path
from statistics import mean

import torch

from ... import Framework
from ...Datasets.utils import View, apply_background_color
from ...Logging import Logger
from .Model import BaseModel


class BaseRenderingComponent(ABC, torch.nn.Module):
    """Sub-component that executes the model.  The reference wraps it in ``torch.nn.DataParallel`` for several
    GPU_INDICES, which cannot scatter a RayBatch (its own FIXME, NeRF/Renderer.py:31); here multi-GPU is one
    process per GPU (nerficg_b200/dist.py), so ``get`` is a plain constructor."""

    @classmethod
    def get(cls, *args) -> 'BaseRenderingComponent':
        return cls(*args)

    @abstractmethod
    def forward(self, *args):
        pass


def quantize_8bit(image: torch.Tensor) -> torch.Tensor:
    """The reference's save_image rounding (Datasets/utils.py:209): clamp * 255 + 0.5 -> uint8."""
    return image.clamp(0.0, 1.0).mul(255.0).add(0.5).to(torch.uint8)


def psnr_8bit(result: torch.Tensor, target: torch.Tensor) -> float:
    a, b = quantize_8bit(result).float() / 255.0, quantize_8bit(target).float() / 255.0
    return float(-10.0 * torch.log10(torch.mean((a - b) ** 2).clamp_min(1e-12)))


class BaseRenderer(Framework.Configurable, ABC):
    def __init__(self, model: BaseModel, valid_model_types: list[type] = None) -> None:
        Framework.Configurable.__init__(self, 'RENDERER')
        ABC.__init__(self)
        if valid_model_types is not None and type(model) not in valid_model_types:
            Logger.log_error(f'provided invalid model for renderer of type: "{type(self)}"\n provided model type: '
                             f'"{type(model)}", valid options are: {valid_model_types}')
            raise Framework.RendererError(f'provided invalid model for renderer of type: "{type(self)}"')
        self.model = model

    @abstractmethod
    def render_image(self, view: View, to_chw: bool = False, benchmark: bool = False) -> dict[str, torch.Tensor | None]:
        pass

    def postprocess_outputs(self, outputs, view, dataset, index) -> dict[str, torch.Tensor]:
        return {'rgb': outputs['rgb'].clamp_(0.0, 1.0)}

    @torch.no_grad()
    def render_subset(self, output_directory: Path | None, dataset, calculate_metrics: bool = False, verbose: bool = True,
                      image_extension: str = 'png', **_) -> dict[str, float]:
        """Renders every view of the dataset's active subset; optionally writes PNGs (torchvision) and returns
        the mean 8-bit PSNR against the ground truth (reference Base/Renderer.py:206-271,103-161)."""
        self.model.eval()
        psnrs = []
        for index, view in enumerate(Logger.log_progress(dataset, total=len(dataset), desc='image', leave=False) if verbose else dataset):
            outputs = self.postprocess_outputs(self.render_image(view, to_chw=True), view, dataset, index)
            if calculate_metrics and view.rgb is not None:
                gt = view.rgb.to(outputs['rgb'].device)
                if view.alpha is not None:
                    gt = apply_background_color(gt, view.alpha.to(gt.device), view.camera.background_color)
                psnrs.append(psnr_8bit(outputs['rgb'], gt))
            if output_directory is not None:
                from torchvision import io
                main = Path(output_directory) / f'{dataset.mode}_{self.model.num_iterations_trained}'
                for key, image in outputs.items():
                    (main / key).mkdir(parents=True, exist_ok=True)
                    io.write_png(quantize_8bit(image).cpu().contiguous(), str(main / key / f'{index:05d}.{image_extension}'), compression_level=6)
        metrics = {'PSNR': mean(psnrs)} if psnrs else {}
        if metrics and output_directory is not None:
            main = Path(output_directory) / f'{dataset.mode}_{self.model.num_iterations_trained}'
            (main / 'metrics_8bit.txt').write_text(f'{self.model.model_name}\nMetric\tMean\nPSNR\t{metrics["PSNR"]:.2f}\n')
        return metrics
