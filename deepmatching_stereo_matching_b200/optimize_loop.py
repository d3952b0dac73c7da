# -*- coding: utf-8 -*-
"""
image_threshold -- the clamp helper used by sub_pix_cal (misc/optimize_loop.py:40-44).
The Gauss-Seidel smoothing loops of that module are a separate post-process and out of
the hot-path scope.
"""

import numpy as np


def image_threshold(arr, threshold=[0, 10]):
    arr = np.where(arr > threshold[1], threshold[1], arr)
    arr = np.where(arr < threshold[0], threshold[0], arr)
    return arr
