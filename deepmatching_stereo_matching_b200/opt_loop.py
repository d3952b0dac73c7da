# -*- coding: utf-8 -*-
"""
Bilateral-weighted Gauss-Seidel loops -- mirror of misc/opt_loop.py:16-85 of the reference (the
`if 0:` branch of optimize_looper.py:55-74): ``make_weight``, ``optimize_loop_bilateral_horizon``,
``optimize_loop_bilateral_vertical``.  The sweeps run on the GPU in the reference's visiting order
(csrc/gauss_seidel.cu): given the same weights they return the reference's bits; the weights of
``make_weight`` use CUDA's exp and agree with numpy's to an ulp.
"""

import numpy as np

from . import _native


def make_weight(guide_img, exclusion, size, sigma):
    """misc/opt_loop.py:66-85 -> (gausian_weight (w, w), color_weight_matrix (S0 - e, S1 - e, w, w))."""
    torch = _native.require_cuda()
    s0, s1, e = int(size[0]), int(size[1]), int(exclusion)
    w = 2 * e + 1
    g = torch.from_numpy(np.ascontiguousarray(guide_img, dtype=np.float64)).cuda()
    if tuple(g.shape) != (s0, s1):
        raise ValueError('guide_img must have shape size = (%d, %d)' % (s0, s1))
    gw = torch.empty((w, w), dtype=torch.float64, device='cuda')
    cw = torch.empty((s0 - e, s1 - e, w, w), dtype=torch.float64, device='cuda')
    _native.check(_native.lib().dm_make_weight(_native.ptr(g), s0, s1, e, float(sigma[0]), float(sigma[1]),
                                               _native.ptr(gw), _native.ptr(cw), _native.stream_ptr()))
    return gw.cpu().numpy(), cw.cpu().numpy()


def _bilateral(img_dis, color_weight_matrix, gausian_weight, coefficient, exclusion, size, vertical):
    torch = _native.require_cuda()
    s0, s1, e = int(size[0]), int(size[1]), int(exclusion)
    w = 2 * e + 1
    d = torch.from_numpy(np.ascontiguousarray(img_dis, dtype=np.float64)).cuda()
    cw = torch.from_numpy(np.ascontiguousarray(color_weight_matrix, dtype=np.float64)).cuda()
    gw = torch.from_numpy(np.ascontiguousarray(gausian_weight, dtype=np.float64)).cuda()
    co = torch.from_numpy(np.ascontiguousarray(coefficient, dtype=np.float64)).cuda()
    if tuple(d.shape) != (s0, s1) or tuple(cw.shape) != (s0 - e, s1 - e, w, w) or tuple(gw.shape) != (w, w):
        raise ValueError('shapes do not match size / exclusion')
    if co.dim() != 2 or co.shape[0] <= e + 1 or co.shape[1] <= e + 1:
        raise IndexError('coefficient too small for exclusion %d' % e)
    # the sweep reads coefficient[e, e], coefficient[e, e +- 1] / [e +- 1, e] only (misc/opt_loop.py:30-31,52-53)
    co = co[:, :].contiguous()
    diff = torch.zeros((s0, s1), dtype=torch.float64, device='cuda')
    err = torch.zeros((1,), dtype=torch.float64, device='cuda')
    # the kernel indexes coefficient with the image's row pitch: hand it a (s0, s1)-pitched copy of the corner it needs
    corner = torch.zeros((s0, s1), dtype=torch.float64, device='cuda')
    corner[:e + 2, :e + 2] = co[:e + 2, :e + 2]
    _native.check(_native.lib().dm_optimize_loop_bilateral(_native.ptr(d), _native.ptr(cw), _native.ptr(gw), _native.ptr(corner), s0, s1, e,
                                                           1 if vertical else 0, _native.ptr(diff), _native.ptr(err), _native.stream_ptr()))
    out = d.cpu().numpy()
    if isinstance(img_dis, np.ndarray) and img_dis.dtype == np.float64 and img_dis.shape == out.shape:
        img_dis[...] = out              # the reference updates its argument in place and returns it
        out = img_dis
    return out, float(err.item())


def optimize_loop_bilateral_horizon(img_dis, color_weight_matrix, gausian_weight, coefficient, alpha, exclusion, size):
    """misc/opt_loop.py:16-38 (alpha is unused there as well)."""
    return _bilateral(img_dis, color_weight_matrix, gausian_weight, coefficient, exclusion, size, False)


def optimize_loop_bilateral_vertical(img_dis, color_weight_matrix, gausian_weight, coefficient, alpha, exclusion, size):
    """misc/opt_loop.py:41-63."""
    return _bilateral(img_dis, color_weight_matrix, gausian_weight, coefficient, exclusion, size, True)
