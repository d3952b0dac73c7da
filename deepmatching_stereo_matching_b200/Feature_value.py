# -*- coding: utf-8 -*-
"""
Feature_value -- patch-vs-image similarity map on the GPU.

Mirror of misc/Feature_value.py:18-43 of the reference: same constructor argument, same
validation (print + sys.exit), same float32 min-maxed result.  The arithmetic of
cv2.matchTemplate (misc/Feature_value.py:41) runs in libdmstereo (dm_feature_value).
"""

import sys

import numpy as np

from . import _native


class Feature_value():

    def __init__(self, feature_name='cv2.TM_CCOEFF_NORMED'):
        FEATURE_NAME_LIST = ['cv2.TM_CCOEFF_NORMED', 'cv2.TM_CCOEFF']
        if feature_name not in FEATURE_NAME_LIST:
            print('invalid feature_name \'{}\' is inputed!'.format(feature_name))
            sys.exit()
        self.feature_name = feature_name
        self.method = _native.method_id(feature_name)          # == the cv2 enum value

    @staticmethod
    def min_max(x, axis=None):
        mn = x.min(axis=axis, keepdims=True)
        mx = x.max(axis=axis, keepdims=True)
        return (x - mn) / (mx - mn)

    def __call__(self, img, template):
        torch = _native.require_cuda()
        a = np.ascontiguousarray(img)
        b = np.ascontiguousarray(template)
        if a.dtype != np.uint8 or b.dtype != np.uint8 or a.ndim != 2 or b.ndim != 2:
            raise TypeError('Feature_value expects two 2-D uint8 arrays')
        # cv2.matchTemplate swaps its arguments when the first one is the smaller
        if a.shape[0] <= b.shape[0] and a.shape[1] <= b.shape[1]:
            patch, image = a, b
        elif b.shape[0] <= a.shape[0] and b.shape[1] <= a.shape[1]:
            patch, image = b, a
        else:
            raise ValueError('neither array fits inside the other: %s vs %s' % (a.shape, b.shape))
        dp = torch.from_numpy(patch).cuda()
        di = torch.from_numpy(image).cuda()
        oh, ow = image.shape[0] - patch.shape[0] + 1, image.shape[1] - patch.shape[1] + 1
        out = torch.empty((oh, ow), dtype=torch.float32, device='cuda')
        _native.check(_native.lib().dm_feature_value(_native.ptr(dp), patch.shape[0], patch.shape[1],
                                                     _native.ptr(di), image.shape[0], image.shape[1],
                                                     self.method, _native.ptr(out), _native.stream_ptr()))
        return out.cpu().numpy()
