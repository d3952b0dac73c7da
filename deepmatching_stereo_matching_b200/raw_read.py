# -*- coding: utf-8 -*-
"""
RawRead -- headerless raw scene reader.  Mirror of misc/raw_read.py:15-45: int8 read,
multiply by ``rate``, wrap to uint8.  Host I/O only.
"""

import numpy as np


class RawRead():

    def __init__(self):
        pass

    @staticmethod
    def _read8(filename, xdata, ydata, band):
        return np.fromfile(filename, dtype=np.int8, count=xdata * ydata * band).reshape(band, ydata, xdata)

    @classmethod
    def read(self, path, size=(6000, 6000), rate=1):
        return (self._read8(path, size[0], size[1], 1) * rate)[0].astype(np.uint8)
