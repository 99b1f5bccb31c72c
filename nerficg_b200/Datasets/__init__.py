from .utils import RayBatch, RayCollection, View, apply_background_color  # noqa: F401
