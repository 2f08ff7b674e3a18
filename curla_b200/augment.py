"""Host side of the two device augmentations (csrc/augment.cu): parameter draws and the ctypes
calls.  kornia (the reference's implementation, augmentations.py:5) is an unpinned dependency
that is not available here; behaviour follows its documentation (see DESIGN.md section 1:
PARITY UNPINNED for the colour transform and the Gaussian noise stream; the cover rows and the
numpy draws of the cover values are exact)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

_OPS = {'brightness': 0, 'contrast': 1, 'saturation': 2, 'hue': 3}


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check(image_batch, aug):
    if not (torch.is_tensor(image_batch) and image_batch.is_cuda and image_batch.dtype == torch.float32):
        raise _lib.CurlaError('%s.training_augmentation expects a float32 CUDA tensor (B, 3*k, H, W)'
                              % type(aug).__name__)
    b, c, h, w = image_batch.shape
    assert c % 3 == 0 and (h, w) == tuple(aug.input_shape), (image_batch.shape, aug.input_shape)
    return image_batch.contiguous(), b * (c // 3), h, w


def color_jiggle(image_batch, aug, params=None, order=None, params_out=None):
    """In place on the batch (the reference also writes through its input, augmentations.py:119).
    params: optional float CUDA tensor [4][B*k] {contrast, saturation, hue, apply}; order: optional
    permutation of (0,1,2,3); by default one random order per call (kornia samples one per forward)."""
    x, n, h, w = _check(image_batch, aug)
    if order is None:
        order = torch.randperm(4).tolist()
    code = sum(int(op) << (4 * k) for k, op in enumerate(order))
    aug._calls = getattr(aug, '_calls', 0) + 1
    if not hasattr(aug, '_seed'):
        aug._seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    with torch.cuda.device(x.device):
        _lib.call('curla_color_jiggle', _lib.ptr(x), n, h, w, _lib.ptr(params), aug._seed, aug._calls,
                  float(aug.contrast), float(aug.saturation), float(aug.hue), float(aug.p), code,
                  _lib.ptr(params_out), _stream(x))
    if x.data_ptr() != image_batch.data_ptr():
        image_batch.copy_(x)
    return image_batch


def noisy_cover(image_batch, aug, noise=None, cover=None):
    """Cover rows + Gaussian noise + clamp, in place.  cover: the three np.random.randint(0, 255)
    values (drawn here in the reference's order, augmentations.py:192-194, when not given)."""
    x, n, h, w = _check(image_batch, aug)
    if cover is None:
        cover = aug.draw_cover()
    cov = (C.c_float * 3)(*[float(v) for v in cover])
    aug._calls = getattr(aug, '_calls', 0) + 1
    if not hasattr(aug, '_seed'):
        aug._seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    if noise is not None:
        noise = noise.to(x.device, torch.float32).contiguous()
    with torch.cuda.device(x.device):
        _lib.call('curla_noisy_cover', _lib.ptr(x), n, h, w, int(aug.top), int(aug.bottom),
                  C.cast(cov, C.c_void_p), float(aug.std), _lib.ptr(noise), aug._seed, aug._calls, _stream(x))
    if x.data_ptr() != image_batch.data_ptr():
        image_batch.copy_(x)
    return image_batch
