from .mel import MelSpectrogramTransform  # noqa: F401
