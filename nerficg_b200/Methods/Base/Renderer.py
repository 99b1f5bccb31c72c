"""Base renderer (reference src/Methods/Base/Renderer.py): model type check, ``render_image`` contract,
test-set rendering with 8-bit PSNR and SSIM, and the online-FPS loop of ``scripts/inference.py -b``.

The reference computes its metrics with torchmetrics (unpinned, absent here -- "parity unpinned" for these two functions):
PSNR = 10 log10(1 / mse) on the 8-bit round trip; SSIM restates torchmetrics' StructuralSimilarityIndexMeasure defaults
(11x11 Gaussian window, sigma 1.5, k1 0.01, k2 0.03, reflect padding cropped again, mean over the image).  LPIPS needs
pretrained VGG weights that cannot be fetched offline and stays out of scope."""
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


def ssim(result: torch.Tensor, target: torch.Tensor, data_range: float = 1.0, sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03) -> float:
    """Structural similarity of two CxHxW (or BxCxHxW) images in [0, data_range] (reference Base/Renderer.py:130-133 uses
    torchmetrics.image.StructuralSimilarityIndexMeasure(data_range=1.0); its defaults are restated here)."""
    a = result[None] if result.dim() == 3 else result
    b = target[None] if target.dim() == 3 else target
    a, b = a.float(), b.float()
    channels = a.shape[1]
    size = int(3.5 * sigma + 0.5) * 2 + 1                      # 11 for sigma = 1.5
    pad = (size - 1) // 2
    x = torch.arange(size, dtype=torch.float32, device=a.device) - (size - 1) / 2
    g1 = torch.exp(-(x / sigma) ** 2 / 2)
    g1 = g1 / g1.sum()
    kernel = (g1[:, None] * g1[None, :]).expand(channels, 1, size, size).contiguous()
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    a = torch.nn.functional.pad(a, (pad, pad, pad, pad), mode='reflect')
    b = torch.nn.functional.pad(b, (pad, pad, pad, pad), mode='reflect')
    stack = torch.cat((a, b, a * a, b * b, a * b))
    mu_a, mu_b, aa, bb, ab = torch.nn.functional.conv2d(stack, kernel, groups=channels).split(a.shape[0])
    var_a, var_b, cov = (aa - mu_a * mu_a).clamp_min(0.0), (bb - mu_b * mu_b).clamp_min(0.0), ab - mu_a * mu_b
    full = ((2 * mu_a * mu_b + c1) * (2 * cov + c2)) / ((mu_a * mu_a + mu_b * mu_b + c1) * (var_a + var_b + c2))
    return float(full[..., pad:-pad, pad:-pad].mean())


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
        psnrs, ssims = [], []
        for index, view in enumerate(Logger.log_progress(dataset, total=len(dataset), desc='image', leave=False) if verbose else dataset):
            outputs = self.postprocess_outputs(self.render_image(view, to_chw=True), view, dataset, index)
            if calculate_metrics and view.rgb is not None:
                gt = view.rgb.to(outputs['rgb'].device)
                if view.alpha is not None:
                    gt = apply_background_color(gt, view.alpha.to(gt.device), view.camera.background_color)
                psnrs.append(psnr_8bit(outputs['rgb'], gt))
                ssims.append(ssim(quantize_8bit(outputs['rgb']).float() / 255.0, quantize_8bit(gt).float() / 255.0))
            if output_directory is not None:
                from torchvision import io
                main = Path(output_directory) / f'{dataset.mode}_{self.model.num_iterations_trained}'
                for key, image in outputs.items():
                    (main / key).mkdir(parents=True, exist_ok=True)
                    io.write_png(quantize_8bit(image).cpu().contiguous(), str(main / key / f'{index:05d}.{image_extension}'), compression_level=6)
        metrics = {'PSNR': mean(psnrs), 'SSIM': mean(ssims)} if psnrs else {}
        if metrics and output_directory is not None:
            main = Path(output_directory) / f'{dataset.mode}_{self.model.num_iterations_trained}'
            (main / 'metrics_8bit.txt').write_text(f'{self.model.model_name}\nMetric\tMean\nPSNR\t{metrics["PSNR"]:.2f}\nSSIM\t{metrics["SSIM"]:.3f}\n')
        return metrics

    @torch.no_grad()
    def benchmark_fps(self, dataset, num_iterations: int = 100, output_path: Path | None = None) -> dict[str, float]:
        """Online FPS of the active subset, the loop of the reference's ``scripts/inference.py -b`` (:62-96): one warm-up pass
        over the views, then ``num_iterations`` passes of ``render_image(view, benchmark=True)`` between two device
        synchronisations, wall-clock timed; writes the reference's ``performance_<iterations>.txt`` when a path is given."""
        from time import perf_counter
        self.model.eval()
        n_views = len(dataset)
        if n_views == 0:
            raise Framework.RendererError('No images found for benchmarking.')
        for view in dataset:
            self.render_image(view, benchmark=True)
        torch.cuda.synchronize()
        start = perf_counter()
        for _ in range(num_iterations):
            for view in dataset:
                self.render_image(view, benchmark=True)
        torch.cuda.synchronize()
        total = perf_counter() - start
        n_images = num_iterations * n_views
        out = {'fps': n_images / total, 'ms_per_image': 1e3 * total / n_images, 'images': n_images, 'total_ms': 1e3 * total}
        if output_path is not None:
            Path(output_path).write_text(
                f'Number of test set renders: {num_iterations}\nNumber of test set images: {n_views}\n'
                f'Test set image size: {view.camera.width}x{view.camera.height}\nTotal rendering time: {out["total_ms"]:.2f} ms\n'
                f'Average rendering time per image: {out["ms_per_image"]:.2f} ms\nAverage FPS: {out["fps"]:.2f}\n')
        return out
