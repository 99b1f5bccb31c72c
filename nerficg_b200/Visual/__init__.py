from .utils import apply_color_map, spectral_lut  # noqa: F401
