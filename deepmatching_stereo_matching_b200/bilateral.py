# -*- coding: utf-8 -*-
"""
bilateral_filter -- cv2.bilateralFilter for 8-bit planes on the GPU: the live branch of the
reference's post-process (optimize_looper.py:76-77 filters the uint8 casts of the two
disparity planes with d = 2*exclusion+1 and the two sigmas).  Same argument order as cv2.
"""

import numpy as np

from . import _native


def bilateral_filter(src, d, sigma_color, sigma_space):
    """uint8 (h,w) numpy array -> uint8 (h,w); OpenCV's own 8-bit algorithm (bit-identical to
    cv2.bilateralFilter with IPP off; IPP builds differ from that by one grey level)."""
    torch = _native.require_cuda()
    a = np.ascontiguousarray(src)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError('bilateral_filter takes a 2-D uint8 array (optimize_looper.py:76 casts with astype(\'uint8\'))')
    s = torch.from_numpy(a).cuda()
    out = torch.empty_like(s)
    _native.check(_native.lib().dm_bilateral_u8(_native.ptr(s), a.shape[0], a.shape[1], int(d), float(sigma_color),
                                                float(sigma_space), _native.ptr(out), _native.stream_ptr()))
    return out.cpu().numpy()
