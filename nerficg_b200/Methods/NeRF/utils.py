"""Host-side pieces of reference src/Methods/NeRF/utils.py: the encoding module (kept for its buffer and for
state-dict parity) and thin functional wrappers with the reference's signatures around the CUDA kernels."""
from __future__ import annotations

import torch

from ... import Framework, ops
from ...Datasets.utils import RayBatch


class FrequencyEncoding(torch.nn.Module):
    """Holds ``frequency_factors`` (1,1,L) = 2^k (utils.py:15-19).  The encoding itself is fused into the first
    layer of the MLP kernel (mlp_fwd.cu); ``forward`` is intentionally not provided on the host."""

    def __init__(self, n_inputs: int, append_input: bool):
        super().__init__()
        self.register_buffer('frequency_factors', torch.linspace(0.0, n_inputs - 1.0, n_inputs).exp2()[None, None, :])
        self.append_input = append_input

    def get_n_outputs(self, n_inputs: int) -> int:
        return n_inputs * 2 * self.frequency_factors.numel() + (n_inputs if self.append_input else 0)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        raise Framework.ModelError('the frequency encoding runs inside the fused CUDA MLP kernel; there is no host path')


SUPPORTED_ACTIVATIONS = ('relu',)


def generate_samples(rays: RayBatch, n_samples: int, near_plane: float, far_plane: float, randomize_samples: bool) -> torch.Tensor:
    """utils.py:57-75 -- draws ``torch.rand((n, n_samples))`` exactly like the reference when randomised."""
    u = torch.rand((len(rays), n_samples), dtype=rays.dtype, device=rays.device) if randomize_samples else None
    return ops.sample_stratified(len(rays), n_samples, near_plane, far_plane, u, rays.device)


def generate_samples_from_pdf(bins: torch.Tensor, values: torch.Tensor, n_samples: int, randomize_samples: bool) -> torch.Tensor:
    """utils.py:78-109 -- returns the (unsorted) fine samples."""
    u = torch.rand(bins.shape[0], n_samples, device=bins.device) if randomize_samples else None
    return ops.sample_importance(bins, values, n_samples, u, return_fine=True)[1]


def integrate_samples(depth_samples, ray_directions, densities, colors, background_color, final_delta: float = 1.0e10):
    """utils.py:112-136 (forward only; the differentiable path is the fused renderer)."""
    if final_delta != 1.0e10:
        raise Framework.RendererError('the CUDA compositing kernel fixes the final interval at 1e10 like the reference default')
    n, s = depth_samples.shape
    rs = torch.cat((colors.reshape(n, s, 3), densities.reshape(n, s, 1)), dim=-1)
    rgb, depth, alpha, w = ops.composite_forward(depth_samples, rs, ray_directions, background_color, want_weights=True)
    return rgb, depth, alpha, w
