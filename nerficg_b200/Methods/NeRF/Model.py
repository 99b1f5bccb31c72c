"""NeRF model with the reference's module tree and state-dict keys (src/Methods/NeRF/Model.py), whose
parameters live in ONE flat fp32 buffer per block so the CUDA kernels, Adam and NCCL each see a single pointer."""
from __future__ import annotations

import torch

from ... import Framework, ops, params
from ..Base.Model import BaseModel
from .utils import SUPPORTED_ACTIVATIONS, FrequencyEncoding


class NeRFBlock(torch.nn.Module):
    """position, direction -> density, colour.  Same constructor, submodule names and parameter shapes as the
    reference block (Model.py:13-57).  The CUDA path is specialised for the shipped architecture; any other
    combination raises (there is no fallback)."""

    def __init__(self, n_layers: int, n_color_layers: int, n_features: int, n_frequencies_position: int,
                 n_frequencies_direction: int, encoding_append_input: bool, input_skips: list[int],
                 activation_function: str) -> None:
        super().__init__()
        if activation_function not in SUPPORTED_ACTIVATIONS:
            raise Framework.ModelError(f'Invalid activation function "{activation_function}" (CUDA path supports {SUPPORTED_ACTIVATIONS})')
        arch = (n_layers, n_color_layers, n_features, n_frequencies_position, n_frequencies_direction,
                bool(encoding_append_input), list(input_skips))
        if arch != (8, 1, 256, 10, 4, True, [5]):
            raise Framework.ModelError(
                f'the sm_100a NeRF kernels are specialised for N_LAYERS=8, N_COLOR_LAYERS=1, N_FEATURES=256, '
                f'N_FREQUENCIES_POSITION=10, N_FREQUENCIES_DIRECTION=4, ENCODING_APPEND_INPUT=True, INPUT_SKIPS=[5]; got {arch}')
        self.input_skips = input_skips
        self.encoding_position = FrequencyEncoding(n_frequencies_position, encoding_append_input)
        self.encoding_direction = FrequencyEncoding(n_frequencies_direction, encoding_append_input)
        n_pos = self.encoding_position.get_n_outputs(3)
        n_dir = self.encoding_direction.get_n_outputs(3)
        relu = lambda: torch.nn.ReLU(True)
        # constructed in the reference's order so the same seed yields the same initial weights
        layers = [torch.nn.Sequential(torch.nn.Linear(n_pos, n_features, bias=True), relu())]
        for i in range(1, n_layers):
            layers.append(torch.nn.Sequential(
                torch.nn.Linear(n_features + (n_pos if i in input_skips else 0), n_features, bias=True), relu()))
        self.initial_layers = torch.nn.ModuleList(layers)
        self.feature_layer = torch.nn.Linear(n_features, n_features, bias=True)
        self.density_layer = torch.nn.Linear(n_features, 1, bias=True)
        self.density_activation = relu()
        self.color_layers = torch.nn.Sequential(torch.nn.Linear(n_features + n_dir, n_features // 2, bias=True), relu(),
                                                torch.nn.Linear(n_features // 2, 3, bias=True), torch.nn.Sigmoid())
        self._flat: torch.Tensor | None = None
        self._alias()

    # ---- flat storage ----------------------------------------------------------------------
    def _named(self) -> dict[str, torch.nn.Parameter]:
        named = dict(self.named_parameters())
        return {name: named[name] for name, _ in params.TENSOR_SPECS}

    def ordered_parameters(self) -> list[torch.nn.Parameter]:
        """The 24 parameters in flat-buffer order."""
        return list(self._named().values())

    def _alias(self) -> None:
        """(Re)creates the flat buffer on the parameters' device and makes every parameter a view into it."""
        named = self._named()
        device = next(iter(named.values())).device
        flat = torch.zeros(params.layout()[2], dtype=torch.float32, device=device)
        with torch.no_grad():
            for name, view in params.views(flat).items():
                view.copy_(named[name].data)
                named[name].data = view
        self._flat = flat

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._alias()
        return self

    @property
    def flat_params(self) -> torch.Tensor:
        """The block's parameters as one contiguous fp32 tensor (padding elements are zero)."""
        named = self._named()
        first = named[params.TENSOR_SPECS[0][0]]
        if self._flat is None or first.data_ptr() != self._flat.data_ptr():
            self._alias()
        return self._flat

    def forward(self, positions: torch.Tensor, directions: torch.Tensor, random_noise_density: float = 0.0):
        """Inference-only convenience with the reference signature (Model.py:59-83): (N,3),(N,3) -> (N,1),(N,3)."""
        flat = self.flat_params
        n = positions.shape[0]
        noise = random_noise_density * torch.randn(n, device=positions.device) if random_noise_density > 0.0 else None
        with torch.no_grad():
            packed = ops.mlp_pack(flat, with_backward=False)
            out = ops.mlp_forward(packed, flat, positions, torch.zeros_like(positions), directions,
                                  torch.zeros(n, 1, device=positions.device), noise).reshape(n, 4)
        return out[:, 3:], out[:, :3]


@Framework.Configurable.configure(
    HIERARCHICAL=True,
    N_LAYERS=8,
    N_COLOR_LAYERS=1,
    N_FEATURES=256,
    N_FREQUENCIES_POSITION=10,
    N_FREQUENCIES_DIRECTION=4,
    ENCODING_APPEND_INPUT=True,
    INPUT_SKIPS=[5],
    NETWORK_ACTIVATION='relu',
)
class NeRF(BaseModel):
    """Coarse + fine NeRFBlock (Model.py:86-128)."""

    def __init__(self, name: str = None) -> None:
        super().__init__(name)
        self.coarse_nerf: NeRFBlock | None = None
        self.nerf: NeRFBlock | None = None

    def _block(self) -> NeRFBlock:
        return NeRFBlock(self.N_LAYERS, self.N_COLOR_LAYERS, self.N_FEATURES, self.N_FREQUENCIES_POSITION,
                         self.N_FREQUENCIES_DIRECTION, self.ENCODING_APPEND_INPUT, self.INPUT_SKIPS, self.NETWORK_ACTIVATION)

    def build(self) -> 'NeRF':
        if self.HIERARCHICAL:
            self.coarse_nerf = self._block()
        self.nerf = self._block()
        return self

    def blocks(self) -> list[NeRFBlock]:
        return [b for b in (self.coarse_nerf, self.nerf) if b is not None]
