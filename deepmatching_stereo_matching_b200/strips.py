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
    """Contiguous [lo,hi) tile-row ranges, sizes differing by at most one (66 -> 9,9,8,...)."""
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
    """Gathers the per-rank row strips of the (planes, rows, width) mosaic on rank ``dst``.

    local: torch tensor (planes, out_h, out_w) on this rank's device; only
    rows row_ranges[rank] are meaningful.  Returns the assembled tensor on ``dst`` and
    None elsewhere.  Strips are padded to the tallest one so a single gather moves
    everything (NCCL: NVLink/NVSwitch; payload is a few hundred MB at most).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    tall = max(hi - lo for lo, hi in row_ranges)
    planes, out_h, out_w = local.shape
    lo, hi = row_ranges[rank]
    send = torch.zeros((planes, tall, out_w), dtype=local.dtype, device=local.device)
    if hi > lo:
        send[:, :hi - lo] = local[:, lo:hi]
    # only `dst` needs the strips: a gather moves (world-1) strips into one GPU instead of
    # world*(world-1) for an all-gather
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.empty((planes, out_h, out_w), dtype=local.dtype, device=local.device)
    for r, (a, b) in enumerate(row_ranges):
        if b > a:
            full[:, a:b] = recv[r][:, :b - a]
    return full


class SharedHostMosaic(object):
    """The whole (planes, out_h, out_w) mosaic in ONE host buffer shared by the ranks of a node
    (POSIX shared memory, page-locked in every process that has a CUDA device).  Every rank
    copies its own strip device -> host over its own PCIe link with ``copy_strip``; after a
    barrier the assembled mosaic is visible to all ranks as ``.array`` -- the host-side
    alternative to gathering the strips on one GPU and reading 8 strips through one link."""

    def __init__(self, shape, dtype=np.float64, group=None):
        from multiprocessing import shared_memory, resource_tracker
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        name = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
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

    def close(self):
        import torch
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
            self._registered = False
        self.array = None
        self._shm.close()
        if self.rank == 0:
            self._shm.unlink()


class StripSolver(object):
    """Device-resident solve of this rank's strip + gather.  ``solve()`` returns the
    (n_modes+1, out_h, out_w) float64 mosaic (disparity planes, then the score plane) on
    rank 0 and None on the other ranks."""

    def __init__(self, shape, image_size, stride, window_size, feature_name='cv2.TM_CCOEFF_NORMED',
                 degree_map_mode=('elevation',), sub_pix=True, group=None, fused=-1):
        import torch.distributed as dist
        from . import _native
        self._native = _native
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        full = _native.scene_geometry(_native.scene_params(shape, image_size, stride, window_size, feature_name,
                                                           list(degree_map_mode), sub_pix))
        self.len0, self.len1, self.out_h, self.out_w = full.len0, full.len1, full.out_h, full.out_w
        self.parts = partition_tile_rows(full.len0, self.world)
        self.row_ranges = [strip_rows(lo, hi, full.len0, stride[0], image_size[0]) if hi > lo else (0, 0) for lo, hi in self.parts]
        lo, hi = self.parts[self.rank]
        self.tile_rows = (lo, hi)
        self.prm = _native.scene_params(shape, image_size, stride, window_size, feature_name, list(degree_map_mode),
                                        sub_pix, (lo, hi), fused) if hi > lo else None
        self.n_planes = len(degree_map_mode) + 1
        self.ctx = _native.Context()
        self.info = None

    def alloc_planes(self):
        torch = self._native.require_cuda()
        return torch.zeros((self.n_planes, self.out_h, self.out_w), dtype=torch.float64, device='cuda')

    def solve_local(self, img1_dev, img2_dev, planes):
        if self.prm is not None:
            self.info = self.ctx.solve_device(self.prm, img1_dev, img2_dev, planes[:-1], planes[-1])
        return planes

    def gather(self, planes):
        if self.world == 1:
            return planes
        return gather_strips(planes, self.row_ranges, self.group)
