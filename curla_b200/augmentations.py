"""Drop-in for the reference's augmentations.py (augmentations.py:7-221).

Same classes, constructor arguments, attributes and method names.  The training-time
work runs in CUDA kernels (gather.cu / augment.cu); the numpy global RNG is consumed in
exactly the reference's order so that sampled crops are bit-identical
(augmentations.py:63-67).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _as_device_u8(image_batch, device):
    if isinstance(image_batch, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(image_batch)).to(device)
    return image_batch.contiguous()


class IdentityAugmentation:
    def __init__(self, input_shape):
        assert len(input_shape) == 2, "Input shape must be 2D"
        self.input_shape = input_shape
        self.output_shape = input_shape

    def evaluation_augmentation(self, image):
        return image

    def training_augmentation(self, image_batch):
        return image_batch


class RandomCrop(IdentityAugmentation):
    def __init__(self, input_shape):
        super().__init__(input_shape)
        self.cropping_factor = 0.84
        self.output_shape = tuple(int(np.ceil(x * self.cropping_factor)) for x in self.input_shape)

    def evaluation_augmentation(self, image):
        """Centre crop (augmentations.py:26-45); plain slicing of one CHW image."""
        h, w = self.input_shape
        new_h, new_w = self.output_shape
        top = (h - new_h) // 2
        left = (w - new_w) // 2
        return image[:, top:top + new_h, left:left + new_w]

    def draw_offsets(self, n, img_shape=None):
        """h1 then w1 from the numpy GLOBAL stream, exclusive upper bound (in - out):
        augmentations.py:63-67."""
        img_shape = img_shape or self.input_shape
        crop_max_h = img_shape[0] - self.output_shape[0]
        crop_max_w = img_shape[1] - self.output_shape[1]
        h1 = np.random.randint(0, crop_max_h, n)
        w1 = np.random.randint(0, crop_max_w, n)
        return h1, w1

    def training_augmentation(self, image_batch, device=None):
        """(B, C, H, W) uint8 numpy array or CUDA tensor -> randomly cropped batch of the
        same kind (augmentations.py:47-75), cropped by the K1 kernel."""
        was_numpy = isinstance(image_batch, np.ndarray)
        if device is None:
            device = image_batch.device if not was_numpy else torch.device('cuda')
        n = image_batch.shape[0]
        h1, w1 = self.draw_offsets(n, tuple(image_batch.shape[2:4]))
        x = _as_device_u8(image_batch, device)
        assert x.dtype == torch.uint8, 'RandomCrop.training_augmentation expects uint8 frames'
        off = torch.from_numpy(np.stack([h1, w1]).astype(np.int64)).to(device)
        oh, ow = self.output_shape
        out = torch.empty((n, x.shape[1], oh, ow), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.call('curla_gather_crop_f32', _lib.ptr(x), x.shape[1], x.shape[2], x.shape[3], None,
                      _lib.ptr(off[0]), _lib.ptr(off[1]), n, oh, ow, _lib.ptr(out),
                      C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        out = out.to(torch.uint8)
        return out.cpu().numpy() if was_numpy else out


class ColorJiggle(IdentityAugmentation):
    """kornia ColorJiggle(brightness 0, contrast 0.2, saturation 0.5, hue 0.5, p 0.85) on
    every RGB frame of the stack (augmentations.py:78-136), as one CUDA kernel."""

    def __init__(self, input_shape):
        super().__init__(input_shape)
        self.output_shape = self.input_shape
        self.contrast, self.saturation, self.hue, self.p = 0.2, 0.5, 0.5, 0.85

    def training_augmentation(self, image_batch):
        from . import augment
        return augment.color_jiggle(image_batch, self)


class NoisyCover(IdentityAugmentation):
    """Cover rows [0, ceil(.31h)) and [h-ceil(.2h), h) of every R/G/B plane with one random
    value per channel, add N(0, 10^2), clamp to [0,255] (augmentations.py:138-205)."""

    def __init__(self, input_shape):
        super().__init__(input_shape)
        self.output_shape = self.input_shape
        self.h = self.input_shape[0]
        self.top = int(np.ceil(self.h * 0.31))
        self.bottom = int(np.ceil(self.h * 0.20))
        self.std = 10.0

    def draw_cover(self):
        """three np.random.randint(0, 255) scalars per call: augmentations.py:192-194."""
        return [np.random.randint(0, 255) for _ in range(3)]

    def training_augmentation(self, image_batch):
        from . import augment
        return augment.noisy_cover(image_batch, self)


def make_augmentor(name, input_shape):
    print(f'CHOSEN AUGMENTATION: {name}')
    if name == 'identity':
        return IdentityAugmentation(input_shape)
    elif name == 'random_crop':
        return RandomCrop(input_shape)
    elif name == 'color_jiggle':
        return ColorJiggle(input_shape)
    elif name == 'noisy_cover':
        return NoisyCover(input_shape)
    raise ValueError('augmentation is not supported: %s' % name)
