from .Perspective import PerspectiveCamera, SharedCameraSettings, fov_to_focal, focal_to_fov  # noqa: F401
