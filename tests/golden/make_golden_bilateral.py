# -*- coding: utf-8 -*-
"""
Golden vectors for the live branch of optimize_looper.py (:76-77): cv2.bilateralFilter on the
uint8 cast of a disparity plane.  Build container only (needs cv2):

    python tests/golden/make_golden_bilateral.py

Stored twice: with IPP switched off (OpenCV's own algorithm -- the parity contract, bit exact)
and with the build's default (IPP on in cv2 4.13: small windows differ by one grey level).
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

rng = np.random.default_rng(5)
plane = rng.normal(size=(61, 53)) * 2.5 + 1.0           # a disparity plane: a few pixels either side of zero
plane[10:30, 5:19] += 7.0
plane[40:, 30:] -= 4.0
img = plane.astype('uint8')                              # optimize_looper.py:76 -- negative values wrap
out = {'plane': plane, 'img': img}
for ipp in (False, True):
    cv2.ipp.setUseIPP(ipp)
    for d, sc, ss in ((7, 5, 5), (3, 5, 5), (9, 12.5, 3.0), (-1, 4.0, 2.0)):
        out['out_%s_d%d_%g_%g' % ('ipp' if ipp else 'plain', d, sc, ss)] = cv2.bilateralFilter(img, d, sc, ss)
path = os.path.join(HERE, 'bilateral.npz')
np.savez_compressed(path, **out)
print(path, os.path.getsize(path))
