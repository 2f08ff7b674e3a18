"""Import shim (test infrastructure): placeholders so augmentations.py imports.
kornia itself is unavailable => color_jiggle / noisy_cover parity is unpinned."""


class _Unavailable:
    def __init__(self, *a, **k):
        pass

    def __call__(self, x):
        raise RuntimeError('kornia is not available in this environment')


ColorJiggle = _Unavailable
RandomGaussianNoise = _Unavailable
