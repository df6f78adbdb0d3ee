"""CPU restatement of the reference's training-sample transform (TEST INFRASTRUCTURE ONLY, like all of oracle/).

The reference builds its low-resolution inputs with torchvision's Resize(BICUBIC) on a PIL image
(torchsr/dataset.py:86-91,118-121), i.e. with Pillow (requirements.txt:3 pins Pillow==9.0.1, setup.py:44 asks
>= 7.1.2; 12.2.0 in this image) - a third-party dependency that is not under /root/reference. Its 8-bit resize
(src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
ImagingResampleVertical_8bpc) is restated here in numpy integer arithmetic and PINNED against Pillow itself in
tests/test_gpu_data.py (bit-exact on random and structured images), so the CUDA kernel can be held to it byte for byte.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x: float, a: float = -0.5) -> float:
    """Resample.c bicubic_filter."""
    x = -x if x < 0.0 else x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coefficients(in_size: int, out_size: int, support: float = 2.0):
    """precompute_coeffs + normalize_coeffs_8bpc: integer taps kk [out][ksize], bounds [out][2] = (first tap, count)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    sup = support * filterscale
    ksize = int(math.ceil(sup)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.int64)
    bounds = np.zeros((out_size, 2), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - sup + 0.5), 0)
        xmax = min(int(center + sup + 0.5), in_size) - xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(v * (1 << PRECISION_BITS) + (-0.5 if v < 0 else 0.5))
        bounds[xx] = (xmin, xmax)
    return kk, bounds


def resize_u8(img: np.ndarray, out_size: int) -> np.ndarray:
    """Image.resize((out, out), BICUBIC) of a square uint8 HxWxC image: horizontal pass, then vertical pass, each
    accumulated in integers from 1 << 21 and clamped to uint8 after >> 22 (ImagingResample*_8bpc, clip8)."""
    H, W, C = img.shape
    kk, bounds = coefficients(W, out_size)
    tmp = np.zeros((H, out_size, C), np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.full((H, C), 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += img[:, xmin + x, :].astype(np.int64) * kk[xx, x]
        tmp[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    kk, bounds = coefficients(H, out_size)
    out = np.zeros((out_size, out_size, C), np.uint8)
    for yy in range(out_size):
        ymin, n = bounds[yy]
        acc = np.full((out_size, C), 1 << (PRECISION_BITS - 1), np.int64)
        for y in range(n):
            acc += tmp[ymin + y].astype(np.int64) * kk[yy, y]
        out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def train_sample(img: np.ndarray, x0: int, y0: int, crop: int, flip_h: bool, flip_v: bool):
    """TrainData.__getitem__ (dataset.py:118-121) for a given crop origin / flips: (lr, hr) float32 CHW in [0, 1]."""
    hr = img[y0:y0 + crop, x0:x0 + crop]
    if flip_h:
        hr = hr[:, ::-1]
    if flip_v:
        hr = hr[::-1]
    hr = np.ascontiguousarray(hr)
    lr = resize_u8(hr, crop // 4)
    to = lambda a: (a.astype(np.float32) / np.float32(255)).transpose(2, 0, 1)  # noqa: E731  (ToTensor)
    return to(lr), to(hr)
