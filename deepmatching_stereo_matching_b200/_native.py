# -*- coding: utf-8 -*-
"""
ctypes binding of libdmstereo.so (include/dmstereo.h).  There is no CPU fallback: if the
library cannot be loaded (or built with nvcc) every entry point of the package raises.
"""

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libdmstereo.so')

TM_CCOEFF = 4
TM_CCOEFF_NORMED = 5
MODE_IDS = {'elevation': 0, 'elevation2': 1, 'distance': 2}
CORR_AUTO, CORR_SIMT, CORR_UMMA = 0, 1, 2
STAGES = ['descriptors', 'correlation', 'normalize', 'aggregate', 'backtrack', 'planes']


class DmError(RuntimeError):
    pass


class SceneParams(Structure):
    _fields_ = [('scene_h', c_int32), ('scene_w', c_int32), ('t0', c_int32), ('t1', c_int32),
                ('s0', c_int32), ('s1', c_int32), ('ws', c_int32), ('method', c_int32),
                ('n_modes', c_int32), ('modes', c_int32 * 4), ('sub_pix', c_int32),
                ('tile_row_lo', c_int32), ('tile_row_hi', c_int32), ('fused', c_int32),
                ('n_scenes', c_int32), ('filter_num', c_int32), ('filter_cfg', c_int32),
                ('tile_lo', c_int32), ('tile_hi', c_int32)]


class SceneInfo(Structure):
    _fields_ = [('len0', c_int32), ('len1', c_int32), ('out_h', c_int32), ('out_w', c_int32),
                ('row_lo', c_int32), ('row_hi', c_int32), ('n_tiles', c_int32), ('levels', c_int32),
                ('n_map', c_int32), ('used_fused', c_int32), ('chunk_tiles', c_int32),
                ('kernel_launches', c_int32)]


# name -> (restype, argtypes); every symbol include/dmstereo.h declares
SIGNATURES = {
    'dm_version': (c_int, []),
    'dm_last_error': (c_char_p, []),
    'dm_device_cc': (c_int, []),
    'dm_kpad': (c_int, [c_int]),
    'dm_descriptors': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'dm_correlation': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'dm_feature_value': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'dm_minmax_rectify': (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'dm_aggregate': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'dm_backtrack_top': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'dm_backtrack_level': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'dm_row_argmax': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'dm_match_filter': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'dm_match_map': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    'dm_cal_map': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'dm_sub_pix_cal': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p]),
    'dm_optimize_loop': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p]),
    'dm_make_weight': (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    'dm_optimize_loop_bilateral': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'dm_sub_pix_cal_host_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, POINTER(c_int32), c_double, c_void_p]),
    'dm_bilateral_u8': (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_double, c_void_p, c_void_p]),
    'dm_ctx_create': (c_int, [POINTER(c_void_p)]),
    'dm_ctx_destroy': (None, [c_void_p]),
    'dm_ctx_set_stream': (c_int, [c_void_p, c_void_p]),
    'dm_correlation_set_pair_mode': (c_int, [c_int]),
    'dm_ctx_set_workspace_limit': (c_int, [c_void_p, c_size_t]),
    'dm_ctx_workspace_bytes': (c_size_t, [c_void_p]),
    'dm_scene_geometry': (c_int, [POINTER(SceneParams), POINTER(SceneInfo)]),
    'dm_owned_rectangles': (c_int, [POINTER(SceneParams), POINTER(c_int32), POINTER(c_int32)]),
    'dm_solve_scene': (c_int, [c_void_p, POINTER(SceneParams), c_void_p, c_void_p, c_void_p, c_void_p, POINTER(SceneInfo)]),
    'dm_solve_scene_host': (c_int, [c_void_p, POINTER(SceneParams), c_void_p, c_void_p, c_void_p, c_void_p, POINTER(SceneInfo)]),
    'dm_solve_scene_stream': (c_int, [c_void_p, POINTER(SceneParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(SceneInfo)]),
    'dm_multi_create': (c_int, [POINTER(c_int), c_int, POINTER(c_void_p)]),
    'dm_multi_destroy': (None, [c_void_p]),
    'dm_multi_device_count': (c_int, [c_void_p]),
    'dm_multi_set_workspace_limit': (c_int, [c_void_p, c_size_t]),
    'dm_partition_tile_rows': (c_int, [c_int, c_int, POINTER(c_int32), POINTER(c_int32)]),
    'dm_multi_solve_scene_host': (c_int, [c_void_p, POINTER(SceneParams), c_int, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(SceneInfo)]),
    'dm_multi_solve_scene': (c_int, [c_void_p, POINTER(SceneParams), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_int, c_int, POINTER(SceneInfo)]),
    'dm_multi_gather_strips': (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32), c_int]),
    'dm_multi_synchronize': (c_int, [c_void_p]),
    'dm_ipc_alloc': (c_int, [c_size_t, POINTER(c_void_p)]),
    'dm_ipc_free': (c_int, [c_void_p]),
    'dm_ipc_export': (c_int, [c_void_p, c_char_p]),
    'dm_ipc_open': (c_int, [c_char_p, POINTER(c_void_p)]),
    'dm_ipc_close': (c_int, [c_void_p]),
    'dm_ctx_enable_timing': (c_int, [c_void_p, c_int]),
    'dm_ctx_stage_ms': (c_int, [c_void_p, POINTER(c_float), POINTER(c_int)]),
}

_lib = None


def lib():
    """The loaded library; builds it in-tree with nvcc if it is missing.  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        try:
            _build.build()
        except Exception as exc:                                   # no silent fallback
            raise DmError('libdmstereo.so is missing and could not be built: %s' % exc)
    try:
        handle = ctypes.CDLL(LIB_PATH)
    except OSError as exc:
        raise DmError('cannot load %s: %s' % (LIB_PATH, exc))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)                                 # AttributeError if a symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    if os.environ.get('DM_CORR_PAIR') is not None:                 # tuning knob: 0 = no CTA pairs in the tcgen05 correlation
        handle.dm_correlation_set_pair_mode(int(os.environ['DM_CORR_PAIR']))
    return _lib


def check(rc):
    if rc != 0:
        raise DmError('libdmstereo error %d: %s' % (rc, lib().dm_last_error().decode('utf-8', 'replace')))


def require_cuda():
    """torch is plumbing only: device memory, streams, torch.distributed."""
    import torch
    if not torch.cuda.is_available():
        raise DmError('deepmatching_stereo_matching_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(None)


_PINNED_POOL = {}


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory, from a small pool of this process: results of
    the solvers are handed out in such arrays so that the device can write them asynchronously, and
    page-locking a few hundred MB costs tens of milliseconds every time it is done anew.  A buffer
    returns to the pool when the array AND every view derived from it have been garbage collected."""
    import weakref
    import numpy as np
    torch = require_cuda()
    shape = tuple(int(x) for x in shape)
    dt = np.dtype(dtype)
    nbytes = max(1, int(np.prod(shape, dtype=np.int64)) * dt.itemsize)
    key = (nbytes + (1 << 20) - 1) >> 20 << 20                      # pooled by size rounded up to 1 MiB
    pool = _PINNED_POOL.setdefault(key, [])
    t = pool.pop() if pool else torch.empty(key, dtype=torch.uint8, pin_memory=True)
    base = t.numpy()                                                # owns the reference to t; every view keeps it alive

    def give_back(pool=pool, t=t, key=key):
        # keep at most four buffers of a size and 8 GiB in all; the rest goes back to the allocator
        if len(pool) < 4 and sum(k * len(v) for k, v in _PINNED_POOL.items()) + key <= (8 << 30):
            pool.append(t)
    weakref.finalize(base, give_back)
    return base[:nbytes].view(dt).reshape(shape)


def method_id(feature_name):
    return {'cv2.TM_CCOEFF_NORMED': TM_CCOEFF_NORMED, 'cv2.TM_CCOEFF': TM_CCOEFF}[feature_name]


class Context(object):
    """Owns a dm_ctx (workspace + staging buffers) bound to torch's current stream."""

    def __init__(self, workspace_limit=None, timing=False):
        require_cuda()
        self._h = c_void_p()
        check(lib().dm_ctx_create(byref(self._h)))
        if workspace_limit:
            check(lib().dm_ctx_set_workspace_limit(self._h, int(workspace_limit)))
        if timing:
            check(lib().dm_ctx_enable_timing(self._h, 1))
        self.timing = timing

    def close(self):
        if getattr(self, '_h', None):
            lib().dm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bind_stream(self):
        check(lib().dm_ctx_set_stream(self._h, stream_ptr()))

    def solve_device(self, prm, img1, img2, d_map, out_map):
        info = SceneInfo()
        self.bind_stream()
        check(lib().dm_solve_scene(self._h, byref(prm), ptr(img1), ptr(img2), ptr(d_map), ptr(out_map), byref(info)))
        return info

    def solve_host(self, prm, img1_np, img2_np, d_map_np, out_map_np):
        info = SceneInfo()
        self.bind_stream()
        check(lib().dm_solve_scene_host(self._h, byref(prm), c_void_p(img1_np.ctypes.data), c_void_p(img2_np.ctypes.data),
                                        c_void_p(d_map_np.ctypes.data), c_void_p(out_map_np.ctypes.data), byref(info)))
        return info

    def set_workspace_limit(self, nbytes):
        check(lib().dm_ctx_set_workspace_limit(self._h, int(nbytes)))

    def sub_pix_cal_host_batch(self, d_maps_np, co_maps_np, directions, ratio, out_np):
        n, n_planes, s0, s1 = d_maps_np.shape
        dirs = (c_int32 * n_planes)(*[int(d) for d in directions])
        check(lib().dm_sub_pix_cal_host_batch(self._h, c_void_p(d_maps_np.ctypes.data), c_void_p(co_maps_np.ctypes.data), int(n), int(n_planes),
                                              int(s0), int(s1), dirs, float(ratio), c_void_p(out_np.ctypes.data)))

    def stage_ms(self):
        ms = (c_float * len(STAGES))()
        ln = (c_int * len(STAGES))()
        check(lib().dm_ctx_stage_ms(self._h, ms, ln))
        return {s: float(ms[i]) for i, s in enumerate(STAGES)}, {s: int(ln[i]) for i, s in enumerate(STAGES)}

    @property
    def workspace_bytes(self):
        return int(lib().dm_ctx_workspace_bytes(self._h))


    def solve_stream(self, prm, img1, img2, d_map, out_map, d_map_dst, out_map_dst):
        """Device-resident solve whose finished rows are also streamed (band by band, on a second
        stream) into d_map_dst / out_map_dst: integers = raw UVA pointers (a peer mosaic opened with
        dm_ipc_open), numpy arrays = page-locked host memory, tensors = device memory."""
        def raw(x):
            if isinstance(x, int):
                return c_void_p(x)
            if hasattr(x, 'data_ptr'):
                return c_void_p(x.data_ptr())
            return c_void_p(x.ctypes.data)
        info = SceneInfo()
        self.bind_stream()
        check(lib().dm_solve_scene_stream(self._h, byref(prm), ptr(img1), ptr(img2), ptr(d_map), ptr(out_map),
                                          raw(d_map_dst), raw(out_map_dst), byref(info)))
        return info


_CTX = {}


def current_context():
    """One dm_ctx (workspace, staging buffers) per CUDA device of this process, for the current device."""
    torch = require_cuda()
    dev = torch.cuda.current_device()
    if dev not in _CTX:
        _CTX[dev] = Context(workspace_limit=WORKSPACE_LIMIT[0])
    return _CTX[dev]


WORKSPACE_LIMIT = [None]
GATHER_P2P, GATHER_NCCL = 0, 1


class MultiContext(object):
    """dm_multi: one dm_ctx, stream and host thread per device of this process."""

    def __init__(self, devices=None, workspace_limit=None):
        require_cuda()
        self._h = c_void_p()
        if devices is None:
            check(lib().dm_multi_create(None, 0, byref(self._h)))
        else:
            arr = (c_int * len(devices))(*[int(d) for d in devices])
            check(lib().dm_multi_create(arr, len(devices), byref(self._h)))
        self.n_devices = int(lib().dm_multi_device_count(self._h))
        if workspace_limit:
            self.set_workspace_limit(workspace_limit)

    def close(self):
        if getattr(self, '_h', None):
            lib().dm_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_workspace_limit(self, nbytes):
        check(lib().dm_multi_set_workspace_limit(self._h, int(nbytes)))

    def solve_host(self, prm, img1_np, img2_np, d_map_np, out_map_np, max_devices=0):
        info = SceneInfo()
        check(lib().dm_multi_solve_scene_host(self._h, byref(prm), int(max_devices), c_void_p(img1_np.ctypes.data), c_void_p(img2_np.ctypes.data),
                                              c_void_p(d_map_np.ctypes.data), c_void_p(out_map_np.ctypes.data), byref(info)))
        return info

    def solve_device(self, prm, imgs1, imgs2, planes, root=0, gather=GATHER_P2P):
        """imgs1 / imgs2 / planes: one tensor per device (uint8 scenes, float64 (n_modes+1, out_h, out_w)).
        Asynchronous: synchronize() before reading planes[root]."""
        n = self.n_devices
        assert len(imgs1) == len(imgs2) == len(planes) == n
        a1 = (c_void_p * n)(*[t.data_ptr() for t in imgs1])
        a2 = (c_void_p * n)(*[t.data_ptr() for t in imgs2])
        pl = (c_void_p * n)(*[t.data_ptr() for t in planes])
        info = SceneInfo()
        check(lib().dm_multi_solve_scene(self._h, byref(prm), a1, a2, pl, int(root), int(gather), byref(info)))
        return info

    def synchronize(self):
        check(lib().dm_multi_synchronize(self._h))


def owned_rectangles(prm):
    """[(row_lo, row_hi, col_lo, col_hi), ...]: the rectangles of the mosaic the tiles of ``prm`` own."""
    rects = (c_int32 * 12)()
    n = c_int32()
    check(lib().dm_owned_rectangles(byref(prm), rects, byref(n)))
    return [tuple(int(rects[4 * k + i]) for i in range(4)) for k in range(n.value)]


def partition_tile_rows(len0, n):
    lo = (c_int32 * n)()
    hi = (c_int32 * n)()
    check(lib().dm_partition_tile_rows(int(len0), int(n), lo, hi))
    return [(int(lo[r]), int(hi[r])) for r in range(n)]


FILTER_IDS = {'median': 0, 'average': 1}


def scene_params(shape, image_size, stride, window_size, feature_name, modes, sub_pix, tile_rows=None, fused=-1, n_scenes=1,
                 filtering=None, tiles=None):
    """filtering = None or (filtering_num, filter_window_size, filtering_mode) of Matching(filtering=True);
    tile_rows = (lo, hi) strip of tile rows, or tiles = (lo, hi) range of tiles in row-major order."""
    prm = SceneParams()
    prm.scene_h, prm.scene_w = int(shape[0]), int(shape[1])
    prm.t0, prm.t1 = int(image_size[0]), int(image_size[1])
    prm.s0, prm.s1 = int(stride[0]), int(stride[1])
    prm.ws = int(window_size)
    prm.method = method_id(feature_name)
    prm.n_modes = len(modes)
    for i, m in enumerate(modes):
        prm.modes[i] = MODE_IDS[m]
    prm.sub_pix = 1 if sub_pix else 0
    prm.tile_row_lo, prm.tile_row_hi = (0, 0) if tile_rows is None else (int(tile_rows[0]), int(tile_rows[1]))
    prm.fused = int(fused)
    prm.n_scenes = int(n_scenes)
    if tiles is not None:
        prm.tile_lo, prm.tile_hi = int(tiles[0]), int(tiles[1])
    if filtering is not None and int(filtering[0]) > 0:
        prm.filter_num = int(filtering[0])
        prm.filter_cfg = int(filtering[1]) | (FILTER_IDS[filtering[2]] << 8)
    return prm


def scene_geometry(prm):
    info = SceneInfo()
    check(lib().dm_scene_geometry(byref(prm), byref(info)))
    return info
