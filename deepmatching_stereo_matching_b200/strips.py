# -*- coding: utf-8 -*-
"""
Multi-GPU partition of a scene: contiguous strips of tile rows, one per rank, solved
independently; one gather of the finished disparity strips to rank 0 (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  SURVEY.md section 8(e); tiles are independent in the
reference (misc/image_cut_solver.py:163-175).

Output rows are owned by exactly one tile row (the covering tile with the largest index,
misc/image_cut_solver.py:165-175 paste order), so the strips are disjoint row ranges of
the mosaic and the gather is a concatenation.
"""

import numpy as np


def partition_tile_rows(len0, world_size):
    """Contiguous [lo,hi) tile-row ranges, sizes differing by at most one (66 -> 9,9,8,...);
    the same rule as dm_partition_tile_rows of the library."""
    base, extra = divmod(int(len0), int(world_size))
    out, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def strip_rows(lo, hi, len0, stride0, image_size0):
    """Output rows [row_lo,row_hi) owned by tile rows [lo,hi)."""
    out_h = stride0 * (len0 - 1) + image_size0
    return stride0 * lo, (out_h if hi == len0 else stride0 * hi)


def input_rows(lo, hi, stride0, image_size0, window_size):
    """Scene rows [a,b) a strip reads: its tiles plus the (T0 + ws - 1 - s0)-row halo."""
    return stride0 * lo, stride0 * (hi - 1) + image_size0 + window_size - 1


def gather_strips(local, row_ranges, group=None, dst=0):
    """Gathers the per-rank row strips of the (planes, rows, width) mosaic on rank ``dst`` with
    point-to-point transfers of the OWNED rows only (NCCL over NVLink on the GPU box, gloo in the
    CPU tests): rank r sends rows row_ranges[r] of every plane, ``dst`` receives them in place.

    local: torch tensor (planes, out_h, out_w) on this rank's device; only rows row_ranges[rank]
    are meaningful.  Returns ``local`` completed to the whole mosaic on ``dst`` and None elsewhere.
    (Rows of a strip are contiguous per plane, so a strip is ``planes`` messages and nothing is
    packed, padded or reassembled.)
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    planes = local.shape[0]
    ops = []
    if rank == dst:
        for r, (a, b) in enumerate(row_ranges):
            if r == dst or b <= a:
                continue
            peer = dist.get_global_rank(group, r) if group is not None else r
            for p in range(planes):
                ops.append(dist.P2POp(dist.irecv, local[p, a:b], peer, group))
    else:
        a, b = row_ranges[rank]
        peer = dist.get_global_rank(group, dst) if group is not None else dst
        for p in range(planes):
            if b > a:
                ops.append(dist.P2POp(dist.isend, local[p, a:b], peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return local if rank == dst else None


class SharedHostMosaic(object):
    """The whole (planes, out_h, out_w) mosaic in ONE host buffer shared by the ranks of a node
    (POSIX shared memory, page-locked in every process that has a CUDA device).  Every rank
    writes its own strip over its own PCIe link -- ``dm_solve_scene_host`` streams the finished
    row bands of the strip straight into ``.array`` -- and announces it with ``publish(step)``;
    ``wait(step)`` returns once every rank has: a flag per rank in the same shared segment, no
    collective and no GPU-side gather (8 strips through rank 0's link)."""

    def __init__(self, shape, dtype=np.float64, group=None):
        from multiprocessing import shared_memory, resource_tracker
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        fbytes = 64 * self.world                       # one cache line per rank's flag
        name = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes + fbytes)
            self._shm.buf[nbytes:nbytes + fbytes] = bytes(fbytes)
            name = [self._shm.name]
        if dist.is_initialized():
            dist.broadcast_object_list(name, src=0, group=group)
        if self.rank != 0:
            self._shm = shared_memory.SharedMemory(name=name[0])
            try:        # the creator unlinks; keep the tracker of this process out of it (Python < 3.13)
                resource_tracker.unregister(self._shm._name, 'shared_memory')
            except Exception:
                pass
        self.array = np.ndarray(shape, dtype=dtype, buffer=self._shm.buf)
        self._flags = np.ndarray((self.world, 8), dtype=np.int64, buffer=self._shm.buf, offset=nbytes)
        self._registered = False
        try:
            import torch
            if torch.cuda.is_available():
                rc = torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, nbytes, 0)
                self._registered = (int(rc) == 0)
        except Exception:
            self._registered = False

    def copy_strip(self, planes, row_range):
        """Asynchronous (when the buffer is page-locked) copy of rows [lo,hi) of every plane."""
        import torch
        lo, hi = row_range
        if hi > lo:
            dst = torch.from_numpy(self.array)
            for p in range(dst.shape[0]):           # rows [lo,hi) of one plane are contiguous: one DMA each
                dst[p, lo:hi].copy_(planes[p, lo:hi], non_blocking=True)

    def publish(self, step):
        """This rank's strip of step ``step`` (>= 1, increasing) is complete in ``.array``."""
        self._flags[self.rank, 0] = step

    def wait(self, step, timeout=60.0):
        """Returns once every rank has published ``step``."""
        import time
        t0 = None
        while int(self._flags[:, 0].min()) < step:
            if t0 is None:
                t0 = time.perf_counter()
            elif time.perf_counter() - t0 > timeout:
                raise RuntimeError('SharedHostMosaic.wait(%d): ranks at %s after %.0f s' % (step, self._flags[:, 0].tolist(), timeout))

    def close(self):
        import torch
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
            self._registered = False
        self.array = None
        self._flags = None
        self._shm.close()
        if self.rank == 0:
            self._shm.unlink()


class _DevicePointer(object):
    """A raw device pointer dressed as a CUDA array so that torch can wrap it without copying."""

    def __init__(self, ptr, shape, typestr='<f8'):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}


class PeerMosaic(object):
    """The whole (planes, out_h, out_w) float64 mosaic in the memory of ONE device (rank ``dst``),
    mapped into every rank of the node through CUDA IPC.  Every rank streams the finished row bands
    of its strip into it over NVLink peer memory WHILE it is still solving (``dm_solve_scene_stream``)
    -- the gather of the strips without a collective after the solve.  ``.tensor`` is the mosaic as
    a torch tensor on ``dst`` (None elsewhere); ``.ptr`` the address in this process."""

    def __init__(self, shape, group=None, dst=0):
        import torch.distributed as dist
        from ctypes import byref, c_void_p, create_string_buffer
        from . import _native
        self._native = _native
        self.shape = tuple(int(x) for x in shape)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dst = dst
        nbytes = int(np.prod(self.shape)) * 8
        handle = [None]
        p = c_void_p()
        if self.rank == dst:
            _native.check(_native.lib().dm_ipc_alloc(nbytes, byref(p)))
            buf = create_string_buffer(64)
            _native.check(_native.lib().dm_ipc_export(p, buf))
            handle = [buf.raw]
        if dist.is_initialized():
            dist.broadcast_object_list(handle, src=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        if self.rank != dst:
            _native.check(_native.lib().dm_ipc_open(handle[0], byref(p)))
        self.ptr = int(p.value)
        self.plane_bytes = self.shape[1] * self.shape[2] * 8
        self.tensor = None
        if self.rank == dst:
            import torch
            self._keep = _DevicePointer(self.ptr, self.shape)
            self.tensor = torch.as_tensor(self._keep, device='cuda')

    def score_ptr(self):
        """address of the last plane (the score mosaic)"""
        return self.ptr + (self.shape[0] - 1) * self.plane_bytes

    def close(self):
        from ctypes import c_void_p
        if self.ptr:
            lib = self._native.lib()
            self.tensor = None
            if self.rank == self.dst:
                self._native.check(lib.dm_ipc_free(c_void_p(self.ptr)))
            else:
                self._native.check(lib.dm_ipc_close(c_void_p(self.ptr)))
            self.ptr = 0


class StripSolver(object):
    """This rank's share of a scene.  Two partitions of the same scene are kept:

    * ``prm``      -- a contiguous range of TILES (row-major), shares differing by at most one tile; used
      wherever the finished pixels are streamed to their destination rectangle by rectangle
      (``solve_into`` a PeerMosaic, ``solve_host_into`` a host mosaic);
    * ``prm_rows`` -- a strip of whole tile rows (66 rows over 8 ranks = 9,9,8,...), whose output is a
      contiguous block of rows: what ``solve_local`` + ``gather`` (NCCL) needs.
    """

    def __init__(self, shape, image_size, stride, window_size, feature_name='cv2.TM_CCOEFF_NORMED',
                 degree_map_mode=('elevation',), sub_pix=True, group=None, fused=-1):
        import torch.distributed as dist
        from . import _native
        self._native = _native
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        args = (shape, image_size, stride, window_size, feature_name, list(degree_map_mode), sub_pix)
        full = _native.scene_geometry(_native.scene_params(*args))
        self.len0, self.len1, self.out_h, self.out_w = full.len0, full.len1, full.out_h, full.out_w
        self.parts = partition_tile_rows(full.len0, self.world)
        self.row_ranges = [strip_rows(lo, hi, full.len0, stride[0], image_size[0]) if hi > lo else (0, 0) for lo, hi in self.parts]
        lo, hi = self.parts[self.rank]
        self.tile_rows = (lo, hi)
        self.prm_rows = _native.scene_params(*args, tile_rows=(lo, hi), fused=fused) if hi > lo else None
        self.tile_parts = partition_tile_rows(full.len0 * full.len1, self.world)
        ta, tb = self.tile_parts[self.rank]
        self.tiles = (ta, tb)
        self.prm = _native.scene_params(*args, fused=fused, tiles=(ta, tb)) if tb > ta else None
        # scene rows either share reads
        rows = []
        if hi > lo:
            rows.append(input_rows(lo, hi, stride[0], image_size[0], window_size))
        if tb > ta:
            rows.append(input_rows(ta // full.len1, (tb - 1) // full.len1 + 1, stride[0], image_size[0], window_size))
        self.input_rows = (min(r[0] for r in rows), max(r[1] for r in rows)) if rows else (0, 0)
        self.n_planes = len(degree_map_mode) + 1
        self.ctx = _native.Context()
        self.info = None

    def alloc_planes(self):
        torch = self._native.require_cuda()
        return torch.zeros((self.n_planes, self.out_h, self.out_w), dtype=torch.float64, device='cuda')

    def solve_local(self, img1_dev, img2_dev, planes):
        """Solves this rank's strip of tile rows into ``planes`` (rows ``row_ranges[rank]``)."""
        if self.prm_rows is not None:
            self.info = self.ctx.solve_device(self.prm_rows, img1_dev, img2_dev, planes[:-1], planes[-1])
        return planes

    def gather(self, planes):
        if self.world == 1:
            return planes
        return gather_strips(planes, self.row_ranges, self.group)

    def solve_into(self, img1_dev, img2_dev, planes, mosaic):
        """Solves this rank's tiles and streams what they own into ``mosaic`` (a PeerMosaic) while the
        rest is still being solved.  The owner of the mosaic solves in place.  Asynchronous on the
        current stream, which also waits for the copies."""
        if self.prm is None:
            return
        if mosaic.tensor is not None:
            self.info = self.ctx.solve_device(self.prm, img1_dev, img2_dev, mosaic.tensor[:-1], mosaic.tensor[-1])
        else:
            self.info = self.ctx.solve_stream(self.prm, img1_dev, img2_dev, planes[:-1], planes[-1], mosaic.ptr, mosaic.score_ptr())

    def solve_host_into(self, img1_host, img2_host, host_mosaic):
        """End to end for this rank's tiles: uploads the scene rows they read from the (page-locked)
        host scenes, solves, and streams the finished pixels into ``host_mosaic`` (a SharedHostMosaic
        or any (n_modes+1, out_h, out_w) float64 array).  Returns when everything has landed."""
        if self.prm is None:
            return
        arr = host_mosaic.array if hasattr(host_mosaic, 'array') else host_mosaic
        self.info = self.ctx.solve_host(self.prm, img1_host, img2_host, arr[:-1], arr[-1])
