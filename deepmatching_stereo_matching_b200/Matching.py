# -*- coding: utf-8 -*-
"""
Matching -- top-down backtracking over the correlation pyramid and the level-0 parabola
refinement, on the GPU.

Mirror of misc/Matching.py:20-268 of the reference: same constructor, same (3,T0,T1)
float64 result.  ``Co_obj`` may be this package's Correlation_map (device-resident
pyramid, float32) or any object with ``co_map_list`` (4-D numpy arrays, float32 or
float64 -- processed in their own dtype, so indices are bit-exact with the reference) and
``N_map``.
"""

import sys

import numpy as np

from . import _native
from .Correlation_map import DeviceLevels


class Zero_padding(object):
    """misc/Matching.py:258-268 -- ZeroPad2d(1); the kernels read zeros outside the map."""

    def eval(self):
        return self

    def forward(self, x):
        import torch
        return torch.nn.functional.pad(x, (1, 1, 1, 1))

    __call__ = forward


class Matching():

    def __init__(
        self,
        Co_obj=None,
        filter_window_size=3,
        filtering=False,
        filtering_num=3,
        filtering_mode='median',
        sub_pix=True
    ):
        try:
            Co_obj.co_map_list
        except AttributeError as e:
            print('Error!: {}'.format(e))
            print('please run \'obj=Correlation_map()\' and \'obj()\' first.')
            sys.exit()

        MODES = ['average', 'median']
        assert filtering_mode in MODES, 'invalid filtering mode is input!: {}'.format(filtering_mode)

        self.obj = Co_obj
        self.Padding = Zero_padding()
        self.Padding.eval()

        self.filtering_num = filtering_num
        self.filter_window_size = filter_window_size
        self.filtering = filtering
        self.filtering_mode = filtering_mode
        self.sub_pix = sub_pix

    # ------------------------------------------------------------------ device levels
    def _device_levels(self):
        torch = _native.require_cuda()
        lst = self.obj.co_map_list
        if isinstance(lst, DeviceLevels):
            return lst.device, False      # float32 on the device; __call__ asks for the float64 parabola (is_f64 = 2)
        out = []
        f64 = any(np.asarray(x).dtype != np.float32 for x in lst)
        for x in lst:
            a = np.ascontiguousarray(x, dtype=np.float64 if f64 else np.float32)
            out.append(torch.from_numpy(a).cuda())
        return out, f64

    def _filter_device(self, match, score, torch):
        """One application of Matching._filter (misc/Matching.py:91-93,136-138) = dm_match_filter.
        The reference sizes its snapshot (shape[1], shape[1]) (:235-236): on a non-square map it
        raises ValueError (operands cannot be broadcast, or round() of the NaN mean of an empty
        window), and so does this class -- there is nothing to accelerate and no host fallback."""
        self.filtering_num -= 1
        _, h, w = match.shape
        if not (h >= self.filter_window_size and w >= self.filter_window_size):
            return match
        if h != w:
            raise ValueError('Matching._filter is undefined on non-square maps (%d x %d): the reference '
                             'fails in misc/Matching.py:235-248' % (h, w))
        out = torch.empty_like(match)
        _native.check(_native.lib().dm_match_filter(_native.ptr(match), 1, h, w, int(self.filter_window_size),
                                                    _native.FILTER_IDS[self.filtering_mode], _native.ptr(out), _native.stream_ptr()))
        return out

    # ------------------------------------------------------------------ reference API
    def __call__(self):
        torch = _native.require_cuda()
        lib = _native.lib()
        levels, f64 = self._device_levels()
        sdt = torch.float64 if f64 else torch.float32
        st = _native.stream_ptr
        top = levels[-1]
        a, b = top.shape[:2]
        match = torch.empty((2, a, b), dtype=torch.int32, device='cuda')
        score = torch.empty((a, b), dtype=sdt, device='cuda')
        _native.check(lib.dm_backtrack_top(_native.ptr(top), int(f64), 1, a, b, _native.ptr(match), _native.ptr(score), st()))
        if self.filtering and self.filtering_num > 0:
            match = self._filter_device(match, score, torch)
        for k in range(len(levels) - 2, -1, -1):
            lv = levels[k]
            A, B, C, D = lv.shape
            nmatch = torch.empty((2, A, B), dtype=torch.int32, device='cuda')
            nscore = torch.empty((A, B), dtype=sdt, device='cuda')
            _native.check(lib.dm_backtrack_level(_native.ptr(lv), int(f64), 1, A, B, C, D, _native.ptr(match),
                                                 _native.ptr(nmatch), _native.ptr(nscore), st()))
            match, score = nmatch, nscore
            if self.filtering and self.filtering_num > 0:
                match = self._filter_device(match, score, torch)
        t0, t1 = levels[0].shape[:2]
        out = torch.empty((3, t0, t1), dtype=torch.float64, device='cuda')
        own = isinstance(self.obj.co_map_list, DeviceLevels)      # the library's own pyramid: the reference sees float64 values
        _native.check(lib.dm_match_map(_native.ptr(levels[0]), 2 if own else int(f64), 1, t0, t1, _native.ptr(match), _native.ptr(score),
                                       1 if self.sub_pix else 0, _native.ptr(out), st()))
        self.map = out.cpu().numpy()
        return self.map
