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


def sub_pix_cal_batch(d_maps, co_maps, directions, ratio=100., chunks=16):
    """sub_pix_cal for a batch: d_maps (n, n_modes, S0, S1), co_maps (n, S0, S1), directions[m] per
    plane -> (n, n_modes, S0, S1) float64; each slice identical to sub_pix_cal(d_maps[b, m],
    co_maps[b], directions[m]).  The batch goes through the device in ``chunks`` pieces on two
    streams, so the upload of one piece, the kernels of the previous one and the download of the one
    before share the PCIe link in both directions (page-locked inputs, e.g. solve_batch's results)."""
    torch = _native.require_cuda()
    a_h = torch.from_numpy(np.ascontiguousarray(d_maps, dtype=np.float64))
    c_h = torch.from_numpy(np.ascontiguousarray(co_maps, dtype=np.float64))
    if a_h.dim() != 4 or c_h.dim() != 3 or a_h.shape[0] != c_h.shape[0] or a_h.shape[2:] != c_h.shape[1:] or len(directions) != a_h.shape[1]:
        raise ValueError('d_maps (n, n_modes, S0, S1), co_maps (n, S0, S1), one direction per plane')
    if any(d not in (0, 1) for d in directions):
        raise ValueError('direction must be 0 or 1')
    n = a_h.shape[0]
    res = torch.from_numpy(_native.pinned_empty(tuple(a_h.shape), np.float64))
    lib = _native.lib()
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    step = max(1, -(-n // max(1, int(chunks))))
    for k, b0 in enumerate(range(0, n, step)):
        b1 = min(n, b0 + step)
        st = streams[k % 2]
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            a = a_h[b0:b1].cuda(non_blocking=True)
            c = c_h[b0:b1].cuda(non_blocking=True)
            out = torch.empty_like(a)
            for b in range(b1 - b0):
                for m in range(a.shape[1]):
                    _native.check(lib.dm_sub_pix_cal(_native.ptr(a[b, m]), _native.ptr(c[b]), a.shape[2], a.shape[3], int(directions[m]),
                                                     float(ratio), _native.ptr(out[b, m]), _native.stream_ptr()))
            res[b0:b1].copy_(out, non_blocking=True)
    for st in streams:
        st.synchronize()
    return res.numpy()
