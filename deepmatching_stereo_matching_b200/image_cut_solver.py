# -*- coding: utf-8 -*-
"""
ImageCutSolver -- cuts a scene into halo'd tiles, matches every tile and mosaics the
disparity planes.  Mirror of misc/image_cut_solver.py:26-197 of the reference (same
constructor, same (d_map, out_map) float64 result, same tile grid and overlap rule).

The reference solves the tiles one after the other on the CPU; here all tiles of the
scene (or of this rank's strip of tile rows) go through libdmstereo in a few batched
launches: descriptors -> tcgen05 correlation -> pyramid -> backtracking -> planes.
"""

import numpy as np

from . import _native
from .Correlation_map import Correlation_map
from .Matching import Matching
from .Calc_difference import Calc_difference

_CTX = _native._CTX            # one dm_ctx per CUDA device of this process (shared with sub_pix_cal_batch)


def _context():
    """One dm_ctx (workspace, staging buffers) per CUDA device of this process."""
    return _native.current_context()


_MULTI = {}
_WS_LIMIT = None
MIN_TILES_PER_DEVICE = 128      # below this a device's share is launch-bound: fewer devices are used


def visible_devices():
    """The devices a solve may use.  DM_DEVICES = 'all' | a count | a comma list of indices; without
    it every visible device -- except under a one-process-per-GPU launcher (WORLD_SIZE > 1), where the
    process keeps to its current device and the strips are spread over the ranks (strips.py)."""
    import os
    torch = _native.require_cuda()
    n = torch.cuda.device_count()
    env = os.environ.get('DM_DEVICES', '').strip()
    if env and env != 'all':
        if ',' in env:
            return [int(x) for x in env.split(',') if x.strip() != '']
        return list(range(min(n, max(1, int(env)))))
    if not env and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        return [torch.cuda.current_device()]
    return list(range(n))


def _multi_context(devices):
    """One dm_multi (a dm_ctx, a stream and a host thread per device) per device set of this process."""
    key = tuple(devices)
    if key not in _MULTI:
        _MULTI[key] = _native.MultiContext(list(devices), workspace_limit=_WS_LIMIT)
    return _MULTI[key]


def solve_host(prm, img1, img2, d_map, out_map, devices=None):
    """dm_solve_scene_host on one device, dm_multi_solve_scene_host when the scene is large enough
    to be worth several: tile rows (or, for a batch, whole pairs) are shared out, the result is
    bit-identical either way.  devices: None = visible_devices(), or a list of device indices."""
    devs = visible_devices() if devices is None else list(devices)
    info = _native.scene_geometry(prm)
    parts = prm.n_scenes if prm.n_scenes > 1 else info.len0
    use = max(1, min(len(devs), parts, info.n_tiles // MIN_TILES_PER_DEVICE)) if devices is None else min(len(devs), parts)
    torch = _native.require_cuda()
    if use <= 1 and (devices is None or devs[0] == torch.cuda.current_device()):
        return _context().solve_host(prm, img1, img2, d_map, out_map)
    return _multi_context(devs).solve_host(prm, img1, img2, d_map, out_map, max_devices=use)


class _TileIndex(object):
    """[[i, j], ...] in the reference's order (j outer, i inner), computed on access."""

    def __init__(self, ln):
        self.len0, self.len1 = max(int(ln[0]), 0), max(int(ln[1]), 0)

    def __len__(self):
        return self.len0 * self.len1

    def __getitem__(self, k):
        n = len(self)
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(n))]
        if k < 0:
            k += n
        if not 0 <= k < n:
            raise IndexError('list index out of range')
        return [k % self.len0, k // self.len0]

    def __iter__(self):
        return (self[k] for k in range(len(self)))


class _TileViews(object):
    """img[s0*i : s0*i + t0, s1*j : s1*j + t1] for every tile of a _TileIndex, computed on access."""

    def __init__(self, img, index, stride, trimmed):
        self.img, self.index, self.stride, self.trimmed = img, index, stride, trimmed

    def __len__(self):
        return len(self.index)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        i, j = self.index[k]
        ys, xs = self.stride[0] * i, self.stride[1] * j
        return self.img[ys:ys + self.trimmed[0], xs:xs + self.trimmed[1]]

    def __iter__(self):
        return (self[k] for k in range(len(self)))


def set_workspace_limit(nbytes):
    """Caps the device workspace of this process's solver context; scenes whose tiles do not
    fit are processed in equal chunks of tiles (default limit: 48 GB)."""
    global _WS_LIMIT
    _WS_LIMIT = int(nbytes)
    _native.WORKSPACE_LIMIT[0] = int(nbytes)
    _context().set_workspace_limit(nbytes)
    for m in _MULTI.values():
        m.set_workspace_limit(nbytes)


pinned_empty = _native.pinned_empty       # page-locked result arrays from a small pool (see _native.pinned_empty)


def solve_batch(imgs1, imgs2, image_size=[32, 32], stride=[32, 32], window_size=5,
                feature_name='cv2.TM_CCOEFF_NORMED', degree_map_mode=['elevation'], sub_pix=True, fused=-1, devices=None):
    """ImageCutSolver(...)() for a batch of equally sized scene pairs in ONE library call
    (BASELINE config 4: 64 pairs of 512x512).  imgs1, imgs2: uint8 (n, S0, S1).
    Returns (d_maps float64 (n, n_modes, S0', S1'), out_maps float64 (n, S0', S1')), each
    slice identical to what ImageCutSolver returns for that pair."""
    a = np.ascontiguousarray(imgs1, dtype=np.uint8)
    b = np.ascontiguousarray(imgs2, dtype=np.uint8)
    assert a.ndim == 3 and a.shape == b.shape, 'two (n, S0, S1) uint8 stacks of the same shape'
    n = a.shape[0]
    prm = _native.scene_params(a.shape[1:], image_size, stride, window_size, feature_name, list(degree_map_mode),
                               sub_pix, None, fused, n_scenes=n)
    info = _native.scene_geometry(prm)
    d_maps = pinned_empty((n, len(degree_map_mode), info.out_h, info.out_w), np.float64)
    out_maps = pinned_empty((n, info.out_h, info.out_w), np.float64)
    solve_host(prm, a, b, d_maps, out_maps, devices)
    return d_maps, out_maps


class ImageCutSolver():

    def __init__(
        self, img1, img2,
        image_size=[32, 32], stride=[32, 32], window_size=5,
        feature_name='cv2.TM_CCOEFF_NORMED', degree_map_mode=['elevation'],
        padding=False,
        sub_pix=True,
        filtering=False,
        filtering_window_size=3,
        filtering_num=3,
        filtering_mode='average'
    ):
        self.img_shape = img1.shape
        assert self.img_shape == img2.shape, '2枚の画像は同じサイズ！'
        self.img1 = img1
        self.img2 = img2
        self.stride = stride
        self.window_size = window_size
        self.degree_map_mode = degree_map_mode
        self.exclusive_pix = int((window_size - 1) / 2)
        self.image_size = image_size
        self.trimed_size = [image_size[i] + 2 * self.exclusive_pix for i in range(2)]
        self.feature_name = feature_name

        if padding:
            self._padding()

        # misc/image_cut_solver.py:62 -- the last fitting tile is dropped (no +1)
        self.len = [int(np.floor((self.img_shape[i] - self.trimed_size[i]) / self.stride[i])) for i in range(2)]

        self.padding = padding
        self.sub_pix = sub_pix
        self.filtering = filtering
        self.filtering_window_size = filtering_window_size
        self.filtering_num = filtering_num
        self.filtering_mode = filtering_mode

        self.log_flg = True
        # extensions (keyword-only use): strip of tile rows for the multi-GPU partition and
        # the engine selector (-1 auto, 0 materialising, 1 fused)
        self.tile_rows = None
        self.fused = -1
        # devices the tile rows are spread over: None = every visible device when the scene is large
        # enough (visible_devices(), MIN_TILES_PER_DEVICE), or an explicit list of device indices
        self.devices = None

    def _padding(self):
        # misc/image_cut_solver.py:73-93 zeroes image 2 and leaves img_shape stale, so every
        # window of image 2 is flat and the whole result is NaN.  Fenced, not accelerated.
        raise NotImplementedError('padding=True is broken in the reference (image_cut_solver.py:85-93) and not supported')

    def _cut_and_pool(self):
        """misc/image_cut_solver.py:95-113 -- tile views, j outer / i inner.  The three lists are
        sequences that build their elements when asked: the batched solver never looks at the
        tiles themselves, and 2 x 225 numpy views cost more host time than its launches."""
        self.img_index = _TileIndex(self.len)
        self.img1_sub = _TileViews(self.img1, self.img_index, self.stride, self.trimed_size)
        self.img2_sub = _TileViews(self.img2, self.img_index, self.stride, self.trimed_size)

    def _solver(self, solve_image, solve_template):
        """misc/image_cut_solver.py:115-142 -- one tile through the class API (device kernels)."""
        co_cls = Correlation_map(solve_image, solve_template, window_size=self.window_size, feature_name=self.feature_name)
        co_cls()
        if self.log_flg:
            print('complete to create multi-level correlation pyramid')
            print('pyramid level: {}, N={}'.format(co_cls.iteration, co_cls.N_map))
            self.log_flg = False
        cls = Matching(co_cls, sub_pix=self.sub_pix, filtering=self.filtering, filter_window_size=self.filtering_window_size,
                       filtering_num=self.filtering_num, filtering_mode=self.filtering_mode)
        out = cls()
        del co_cls
        del cls
        return np.array([Calc_difference.cal_map(out, mode=m) for m in self.degree_map_mode]), out[2, :, :]

    def _params(self):
        filt = (self.filtering_num, self.filtering_window_size, self.filtering_mode) if self.filtering else None
        return _native.scene_params(self.img_shape, self.image_size, self.stride, self.window_size, self.feature_name,
                                    list(self.degree_map_mode), self.sub_pix, self.tile_rows, self.fused, filtering=filt)

    def _execute_matching(self):
        """misc/image_cut_solver.py:144-179 -- all tiles batched on the GPU."""
        size_list = [self.stride[i] * self.img_index[-1][i] + self.image_size[i] for i in range(2)]   # IndexError when no tile fits, as in the reference
        MODES = ['elevation', 'elevation2', 'distance']
        for m in self.degree_map_mode:
            if m not in MODES:
                print('please input valid mode! {} are ok. yours is \'{}\''.format(MODES, m))
                import sys
                sys.exit()
        if self.filtering and not self._filter_on_device():
            return self._execute_matching_per_tile(size_list)
        img1 = np.ascontiguousarray(self.img1, dtype=np.uint8)
        img2 = np.ascontiguousarray(self.img2, dtype=np.uint8)
        prm = self._params()
        self.d_map = pinned_empty([len(self.degree_map_mode)] + size_list, np.float64)
        self.out_map = pinned_empty(size_list, np.float64)
        self.info = solve_host(prm, img1, img2, self.d_map, self.out_map, self.devices)
        if self.log_flg:
            print('complete to create multi-level correlation pyramid')
            print('pyramid level: {}, N={}'.format(self.info.levels, self.info.n_map))
            self.log_flg = False

    def _filter_on_device(self):
        """The batched solver runs Matching._filter (misc/Matching.py:224-255) as a kernel between
        the levels.  The reference's filter is only defined on square patch grids: on the others
        it raises in the first tile, which is what the tile-by-tile path through Matching does."""
        small = self.image_size[0] < self.filtering_window_size or self.image_size[1] < self.filtering_window_size
        return self.filtering_num <= 0 or small or (self.image_size[0] == self.image_size[1] and 1 <= self.filtering_window_size <= 255)

    def _execute_matching_per_tile(self, size_list):
        """Tile-by-tile variant through the class API (the same kernels, one tile at a time); used
        by the tests and for a displacement filter on non-square grids (raises like the reference)."""
        self.d_map = np.empty([len(self.degree_map_mode)] + size_list, dtype=float)
        self.out_map = np.empty(size_list, dtype=float)
        for idx in range(len(self.img_index)):
            i, j = self.img_index[idx]
            ys, xs = self.stride[0] * i, self.stride[1] * j
            d, s = self._solver(self.img1_sub[idx], self.img2_sub[idx])
            self.d_map[:, ys:ys + self.image_size[0], xs:xs + self.image_size[1]] = d
            self.out_map[ys:ys + self.image_size[0], xs:xs + self.image_size[1]] = s

    def __call__(self):
        self._cut_and_pool()
        self._execute_matching()
        return self.d_map, self.out_map

    @staticmethod
    def image_save(path, arr, threshold=[100, 190]):
        """misc/image_cut_solver.py:186-197."""
        from PIL import Image
        arr = np.where(arr > threshold[1], threshold[1], arr)
        arr = np.where(arr < threshold[0], threshold[0], arr)
        Image.fromarray(arr.astype(np.uint8)).save(path)
