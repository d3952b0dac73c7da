# -*- coding: utf-8 -*-
"""
Correlation_map -- atomic patches, 4-D level-0 correlation and the multi-level pyramid,
computed on the GPU and kept there.

Mirror of misc/Correlation_map.py:29-184 of the reference: same constructor, same private
method names (bad_matching.py:62-70 calls two of them), same attributes.  ``co_map`` and
``co_map_list`` are materialised as float64 numpy arrays only when somebody reads them;
``Matching`` consumes the device tensors directly.
"""

import sys

import numpy as np

from . import _native
from .Feature_value import Feature_value


class DeviceLevels(list):
    """``co_map_list``: behaves like the reference's list of 4-D float64 numpy arrays, but
    the data lives on the GPU (``.device``: list of float32 torch tensors) and is copied to
    the host level by level on first access."""

    def __init__(self, tensors):
        super(DeviceLevels, self).__init__([None] * len(tensors))
        self.device = list(tensors)

    def _get(self, k):
        v = list.__getitem__(self, k)
        if v is None:
            v = self.device[k].cpu().numpy().astype(np.float64)
            list.__setitem__(self, k, v)
        return v

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._get(i) for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        return self._get(k)

    def __iter__(self):
        return (self._get(k) for k in range(len(self)))


class DeviceCoMap(np.lib.mixins.NDArrayOperatorsMixin):
    """``co_map``: the reference's 4-D float64 array (misc/Correlation_map.py:69-87), but the data
    lives on the GPU (``.device``: float32 tensor (T0,T1,T0,T1)).  It turns into a host float64 array
    the first time it is used as one (arithmetic, numpy functions, general indexing).  The one access
    pattern of bad_matching.py:68-70 -- ``co_map[i, j, i, :]``, a patch's own map row -- is served from
    P rows of T1 values fetched with one dm_row_argmax launch instead of the P x P map; the argmax of
    every such row is available directly as ``row_argmax()``."""

    def __init__(self, tensor):
        self.device = tensor
        self._host_arr = None
        self._rows = None
        self._arg = None

    shape = property(lambda self: tuple(self.device.shape))
    ndim = property(lambda self: self.device.dim())
    dtype = property(lambda self: np.dtype(np.float64))
    size = property(lambda self: int(self.device.numel()))

    def __len__(self):
        return self.device.shape[0]

    def _host(self):
        if self._host_arr is None:
            self._host_arr = self.device.cpu().numpy().astype(np.float64)
        return self._host_arr

    def _own_rows(self):
        if self._rows is None:
            torch = _native.require_cuda()
            t0, t1 = self.device.shape[:2]
            arg = torch.empty((t0, t1), dtype=torch.int32, device='cuda')
            rows = torch.empty((t0, t1, t1), dtype=torch.float32, device='cuda')
            _native.check(_native.lib().dm_row_argmax(_native.ptr(self.device), 1, t0, t1, _native.ptr(arg), _native.ptr(rows),
                                                      _native.stream_ptr()))
            self._arg = arg.cpu().numpy().astype(np.int64)
            self._rows = rows.cpu().numpy().astype(np.float64)
        return self._rows

    def row_argmax(self):
        """(T0,T1) int64: np.argmax(co_map[i, j, i, :]) for every patch, computed on the device."""
        self._own_rows()
        return self._arg

    def __getitem__(self, idx):
        if (self._host_arr is None and isinstance(idx, tuple) and len(idx) == 4 and idx[3] == slice(None)
                and all(isinstance(k, (int, np.integer)) for k in idx[:3])):
            t0 = self.device.shape[0]
            i, j, k = int(idx[0]), int(idx[1]), int(idx[2])
            if i % t0 == k % t0 and -t0 <= i < t0 and -t0 <= k < t0:
                return self._own_rows()[i, j]
        return self._host()[idx]

    def __array__(self, dtype=None, copy=None):
        a = self._host()
        return a if dtype is None else a.astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        inputs = tuple(x._host() if isinstance(x, DeviceCoMap) else x for x in inputs)
        if 'out' in kwargs:
            kwargs['out'] = tuple(x._host() if isinstance(x, DeviceCoMap) else x for x in kwargs['out'])
        return getattr(ufunc, method)(*inputs, **kwargs)

    def __array_function__(self, func, types, args, kwargs):
        def conv(x):
            if isinstance(x, DeviceCoMap):
                return x._host()
            if isinstance(x, (list, tuple)):
                return type(x)(conv(y) for y in x)
            return x
        return func(*conv(args), **{k: conv(v) for k, v in kwargs.items()})

    def __getattr__(self, name):            # everything else an ndarray has (astype, reshape, max, ...)
        if name.startswith('__'):
            raise AttributeError(name)
        return getattr(self._host(), name)


class Maxpool(object):
    """misc/Correlation_map.py:176-184 -- MaxPool2d(3, 2, padding=1) over the last two axes
    of a numpy / torch array, evaluated by the aggregation kernel's pooling stage."""

    def __init__(self, window=3, stride=2, padding=1):
        assert (window, stride, padding) == (3, 2, 1), 'only the reference configuration (3, 2, 1) exists'

    def eval(self):
        return self

    def forward(self, x):
        import torch
        return torch.nn.functional.max_pool2d(x, 3, 2, padding=1)

    __call__ = forward


class Correlation_map():

    def __init__(self, img, template, window_size=3, feature_name='cv2.TM_CCOEFF_NORMED'):
        if img.shape != template.shape:
            print('use same size images!')
            sys.exit()
        self.img = img
        self.template = template
        self.window_size = window_size
        self.lam = 1.4

        self.exclusive_pix = int((window_size - 1) / 2)
        self.image_size = [x for x in img.shape]

        self.Feature = Feature_value(feature_name=feature_name)
        self.Maxpool = Maxpool()
        self.Maxpool.eval()

        self._dev = {}

    # ------------------------------------------------------------------ helpers
    def _grid(self):
        if self.window_size % 2 == 0:
            # the reference fails here as well (broadcast of a (ws-1)^2 window into ws^2)
            raise ValueError('window_size must be odd (got %d)' % self.window_size)
        t0 = self.image_size[0] - 2 * self.exclusive_pix
        t1 = self.image_size[1] - 2 * self.exclusive_pix
        if t0 < 1 or t1 < 1:
            raise ValueError('image %s too small for window %d' % (self.image_size, self.window_size))
        return t0, t1

    # ------------------------------------------------------------------ reference API
    def _create_atomic_patch(self):
        """misc/Correlation_map.py:51-67 + the bf16 descriptors / window statistics."""
        torch = _native.require_cuda()
        t0, t1 = self._grid()
        ws = self.window_size
        win = np.lib.stride_tricks.sliding_window_view(np.asarray(self.img), (ws, ws))
        self.atomic_patch = np.ascontiguousarray(win).astype(np.uint8)
        lib = _native.lib()
        kpad = lib.dm_kpad(ws)
        P = t0 * t1
        origin = torch.zeros(2, dtype=torch.int32, device='cuda')
        d = {}
        for name, arr in (('1', self.img), ('2', self.template)):
            a = np.ascontiguousarray(arr)
            if a.dtype != np.uint8:
                a = a.astype(np.uint8)
            scene = torch.from_numpy(a).cuda()
            desc = torch.empty((P, kpad), dtype=torch.bfloat16, device='cuda')
            stat = torch.empty((P * 6,), dtype=torch.float32, device='cuda')   # DM_STAT_FLOATS
            _native.check(lib.dm_descriptors(_native.ptr(scene), a.shape[0], a.shape[1], a.shape[1], _native.ptr(origin),
                                             1, t0, t1, ws, int(name), _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
            d['desc' + name], d['stat' + name] = desc, stat
        d['kpad'] = kpad
        self._dev.update(d)

    def _create_simple_initial_co_map(self, engine=_native.CORR_AUTO):
        """misc/Correlation_map.py:69-87: level-0 map, every [i,j] slice min-maxed."""
        torch = _native.require_cuda()
        t0, t1 = self._grid()
        P = t0 * t1
        lib = _native.lib()
        d = self._dev
        raw = torch.empty((P, P), dtype=torch.float32, device='cuda')
        _native.check(lib.dm_correlation(_native.ptr(d['desc1']), _native.ptr(d['stat1']), _native.ptr(d['desc2']),
                                         _native.ptr(d['stat2']), 1, P, d['kpad'], self.window_size, self.Feature.method,
                                         engine, _native.ptr(raw), _native.stream_ptr()))
        rect = torch.empty_like(raw)
        _native.check(lib.dm_minmax_rectify(_native.ptr(raw), P, P, _native.ptr(raw), _native.ptr(rect), None, None,
                                            _native.stream_ptr()))
        d['co_map'] = raw.view(t0, t1, t0, t1)          # normalised, un-rectified
        d['level0'] = rect.view(t0, t1, t0, t1)
        self.__dict__.pop('co_map', None)

    @property
    def co_map(self):
        if 'co_map' not in self.__dict__:
            if 'co_map' not in self._dev:
                raise AttributeError('co_map')
            self.__dict__['co_map'] = DeviceCoMap(self._dev['co_map'])
        return self.__dict__['co_map']

    @co_map.setter
    def co_map(self, value):
        self.__dict__['co_map'] = value

    def _aggregation(self, map):
        """misc/Correlation_map.py:89-130 on a host or device 4-D array (no rectification)."""
        torch = _native.require_cuda()
        t = map if isinstance(map, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(map, dtype=np.float32))
        t = t.to('cuda', torch.float32).contiguous()
        a, b, c, dd = t.shape
        out = torch.empty((a // 2, b // 2, c // 2, dd // 2), dtype=torch.float32, device='cuda')
        _native.check(_native.lib().dm_aggregate(_native.ptr(t), 1, a, b, c, dd, 0, _native.ptr(out), _native.stream_ptr()))
        return out if isinstance(map, torch.Tensor) else out.cpu().numpy().astype(np.float64)

    def _rectification(self, map):
        return map ** self.lam

    def _multi_level_correlation_pyramid(self):
        """misc/Correlation_map.py:132-156, every level on the device."""
        torch = _native.require_cuda()
        lib = _native.lib()
        levels = [self._dev['level0']]
        cur = levels[0]
        N = 1
        iteration = 1
        while N < min(levels[0].shape[:2]):
            a, b, c, dd = cur.shape
            nxt = torch.empty((a // 2, b // 2, c // 2, dd // 2), dtype=torch.float32, device='cuda')
            _native.check(lib.dm_aggregate(_native.ptr(cur), 1, a, b, c, dd, 1, _native.ptr(nxt), _native.stream_ptr()))
            levels.append(nxt)
            cur = nxt
            N *= 2
            iteration += 1
        self.co_map_list = DeviceLevels(levels)
        self.iteration = iteration
        self.N_map = N

    def __call__(self):
        self._create_atomic_patch()
        self._create_simple_initial_co_map()
        self._multi_level_correlation_pyramid()
        return self.co_map_list
