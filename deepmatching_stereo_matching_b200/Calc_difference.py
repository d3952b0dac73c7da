# -*- coding: utf-8 -*-
"""
Calc_difference -- (3,T0,T1) match map -> disparity plane, on the GPU.
Mirror of misc/Calc_difference.py:17-49 of the reference.
"""

import sys

import numpy as np

from . import _native


class Calc_difference():

    def __init__(self):
        pass

    @staticmethod
    def cal_map(map, mode='elevation'):
        MODES = ['elevation', 'elevation2', 'distance']
        if mode not in MODES:
            print('please input valid mode! {} are ok. yours is \'{}\''.format(MODES, mode))
            sys.exit()
        torch = _native.require_cuda()
        m = torch.from_numpy(np.ascontiguousarray(map, dtype=np.float64)).cuda()
        t0, t1 = m.shape[1], m.shape[2]
        out = torch.empty((t0, t1), dtype=torch.float64, device='cuda')
        _native.check(_native.lib().dm_cal_map(_native.ptr(m), t0, t1, _native.MODE_IDS[mode], _native.ptr(out),
                                               _native.stream_ptr()))
        return out.cpu().numpy()
