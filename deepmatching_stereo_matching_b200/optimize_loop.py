# -*- coding: utf-8 -*-
"""
optimize_loop / image_threshold -- mirror of misc/optimize_loop.py:15-44 of the reference.

``image_threshold`` is the clamp helper sub_pix_cal uses.  ``optimize_loop`` is the 4-neighbour
Gauss-Seidel smoothing of a disparity plane (one forward and one "reverse" in-place sweep); it runs
on the GPU in the reference's own visiting order (csrc/gauss_seidel.cu) and returns the reference's bits.
"""

import numpy as np

from . import _native


def optimize_loop(img_dis, coefficient, alpha, exclusion, size):
    """misc/optimize_loop.py:15-37 -> (img_dis, error).  The input is not modified (the reference clamps
    into a new array first)."""
    torch = _native.require_cuda()
    s0, s1 = int(size[0]), int(size[1])
    d = torch.from_numpy(np.ascontiguousarray(img_dis, dtype=np.float64)).cuda()
    co = torch.from_numpy(np.ascontiguousarray(coefficient, dtype=np.float64)).cuda()
    if tuple(d.shape) != (s0, s1) or tuple(co.shape) != (s0, s1):
        raise ValueError('img_dis and coefficient must have shape size = (%d, %d)' % (s0, s1))
    diff = torch.zeros((s0, s1), dtype=torch.float64, device='cuda')
    err = torch.zeros((1,), dtype=torch.float64, device='cuda')
    _native.check(_native.lib().dm_optimize_loop(_native.ptr(d), _native.ptr(co), s0, s1, int(exclusion), float(alpha),
                                                 _native.ptr(diff), _native.ptr(err), _native.stream_ptr()))
    return d.cpu().numpy(), float(err.item())


def image_threshold(arr, threshold=[0, 10]):
    arr = np.where(arr > threshold[1], threshold[1], arr)
    arr = np.where(arr < threshold[0], threshold[0], arr)
    return arr
