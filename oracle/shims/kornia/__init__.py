from . import augmentation  # noqa: F401
