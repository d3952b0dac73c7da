# -*- coding: utf-8 -*-
"""
sub_pix_cal -- post-hoc parabola refinement of a disparity mosaic with the score mosaic,
on the GPU (float64, bit-identical to numpy).  Mirror of misc/sub_pix_cal.py:22-53.
"""

import numpy as np

from . import _native
from .optimize_loop import image_threshold  # noqa: F401  (re-exported like the reference)


def sub_pix_cal(arr, co_map, direction=0, ratio=100.):
    torch = _native.require_cuda()
    a = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).cuda()
    c = torch.from_numpy(np.ascontiguousarray(co_map, dtype=np.float64)).cuda()
    if a.dim() != 2 or a.shape != c.shape:
        raise ValueError('arr and co_map must be 2-D arrays of the same shape')
    if direction not in (0, 1):
        # the reference never moves the index for other values: plus == minus == centre
        raise ValueError('direction must be 0 or 1')
    out = torch.empty_like(a)
    _native.check(_native.lib().dm_sub_pix_cal(_native.ptr(a), _native.ptr(c), a.shape[0], a.shape[1], int(direction),
                                               float(ratio), _native.ptr(out), _native.stream_ptr()))
    return out.cpu().numpy()
