# -*- coding: utf-8 -*-
"""
Synthetic stereo pairs: seeded, numpy only, no flat patches.

image 2 = texture; image 1 = texture warped by a disparity field, so that the patch of
image 1 at column x matches image 2 at column x - d.
"""

import numpy as np


def _blur(a, sigma):
    """Separable Gaussian blur with reflect padding (numpy only)."""
    r = max(1, int(3 * sigma + 0.5))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    p = np.pad(a, ((r, r), (0, 0)), mode='reflect')
    a = sum(k[i] * p[i:i + a.shape[0]] for i in range(2 * r + 1))
    p = np.pad(a, ((0, 0), (r, r)), mode='reflect')
    a = sum(k[i] * p[:, i:i + a.shape[1]] for i in range(2 * r + 1))
    return a


def texture(shape, seed=0, sigma=1.5, plain_noise=False):
    """uint8 texture: Gaussian-blurred uniform noise rescaled to [0,255] plus 1 bit of
    white noise (keeps every window non-flat), or plain uniform noise."""
    rng = np.random.default_rng(seed)
    if plain_noise:
        return rng.integers(0, 256, size=shape, dtype=np.uint8)
    a = _blur(rng.random(shape), sigma)
    a = (a - a.min()) / (a.max() - a.min())
    a = a * 253.0 + rng.integers(0, 3, size=shape)
    return np.clip(np.rint(a), 0, 255).astype(np.uint8)


def stereo_pair(shape, seed=0, mode='shift', amp=3, sigma=1.5, plain_noise=False):
    """-> (img1, img2) uint8 of ``shape``.

    mode 'shift': constant ``amp`` px horizontal shift; 'sine': integer-rounded smooth
    field d(y,x) = amp*sin(2*pi*y/H)*cos(2*pi*x/W).
    """
    h, w = shape
    pad = int(abs(amp)) + 1
    tex = texture((h, w + 2 * pad), seed, sigma, plain_noise)
    img2 = tex[:, pad:pad + w]
    if mode == 'shift':
        img1 = tex[:, pad - amp:pad - amp + w]
    elif mode == 'sine':
        yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing='ij')
        d = np.rint(amp * np.sin(2 * np.pi * yy / h) * np.cos(2 * np.pi * xx / w)).astype(np.int64)
        img1 = tex[yy, xx + pad - d]
    else:
        raise ValueError(mode)
    return np.ascontiguousarray(img1), np.ascontiguousarray(img2)
