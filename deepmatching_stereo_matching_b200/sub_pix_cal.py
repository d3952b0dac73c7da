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


def sub_pix_cal_batch(d_maps, co_maps, directions, ratio=100.):
    """sub_pix_cal for a batch: d_maps (n, n_modes, S0, S1), co_maps (n, S0, S1), directions[m] per
    plane -> (n, n_modes, S0, S1) float64; each slice identical to sub_pix_cal(d_maps[b, m],
    co_maps[b], directions[m]).  One library call (dm_sub_pix_cal_host_batch): the batch crosses the
    device in pieces on two streams, so the upload of one piece, the kernels of the previous one and
    the download of the one before share the PCIe link in both directions (page-locked inputs, e.g.
    solve_batch's results, make the copies asynchronous)."""
    _native.require_cuda()
    a = np.ascontiguousarray(d_maps, dtype=np.float64)
    c = np.ascontiguousarray(co_maps, dtype=np.float64)
    if a.ndim != 4 or c.ndim != 3 or a.shape[0] != c.shape[0] or a.shape[2:] != c.shape[1:] or len(directions) != a.shape[1]:
        raise ValueError('d_maps (n, n_modes, S0, S1), co_maps (n, S0, S1), one direction per plane')
    if any(d not in (0, 1) for d in directions):
        raise ValueError('direction must be 0 or 1')
    res = _native.pinned_empty(a.shape, np.float64)
    _native.current_context().sub_pix_cal_host_batch(a, c, directions, ratio, res)
    return res
