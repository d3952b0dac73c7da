"""world_size-2 gloo test of the multi-GPU strip path on CPU: each rank solves its strip
of tile rows with the oracle (standing in for the CUDA solve), the strips are gathered
with the production gather code, and rank 0 must hold exactly the whole-scene mosaic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from oracle import dm_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from deepmatching_stereo_matching_b200.strips import partition_tile_rows, strip_rows, gather_strips
        g = load_golden('solver_96_t16_s12_ws5')
        size, stride, ws = tuple(int(x) for x in g['image_size']), tuple(int(x) for x in g['stride']), int(g['ws'])
        modes = tuple(str(m) for m in g['modes'])
        ln, _ = O.tile_grid(g['img1'].shape, size, stride, ws)
        parts = partition_tile_rows(ln[0], world)
        ranges = [strip_rows(lo, hi, ln[0], stride[0], size[0]) for lo, hi in parts]
        lo, hi = parts[rank]
        d, s = O.image_cut_solver(g['img1'], g['img2'], size, stride, ws, modes, bool(g['sub_pix']), tile_rows=(lo, hi))
        local = torch.from_numpy(np.concatenate([d, s[None]], 0))
        a, b = ranges[rank]
        local[:, :a] = float('nan')          # rows this rank does not own must not matter
        local[:, b:] = float('nan')
        full = gather_strips(local, ranges)
        # host-side assembly: every rank writes its strip into one shared host buffer
        from deepmatching_stereo_matching_b200.strips import SharedHostMosaic
        mosaic = SharedHostMosaic(tuple(local.shape), np.float64)
        mosaic.copy_strip(local, ranges[rank])
        mosaic.publish(1)                    # a flag per rank in the shared segment instead of a collective
        mosaic.wait(1)
        shared = mosaic.array.copy()
        dist.barrier()
        mosaic.close()
        # the tile-granular shares of the streaming paths: this rank's range of tiles, pasted through the
        # rectangles it owns (dm_owned_rectangles) into a second shared mosaic
        from deepmatching_stereo_matching_b200 import _native
        args = (g['img1'].shape, size, stride, ws, 'cv2.TM_CCOEFF_NORMED', list(modes), bool(g['sub_pix']))
        tparts = partition_tile_rows(ln[0] * ln[1], world)
        ta, tb = tparts[rank]
        d2, s2 = O.image_cut_solver(g['img1'], g['img2'], size, stride, ws, modes, bool(g['sub_pix']), tiles=(ta, tb))
        mine = np.concatenate([d2, s2[None]], 0)
        mosaic2 = SharedHostMosaic(tuple(local.shape), np.float64)
        for r0, r1, c0, c1 in _native.owned_rectangles(_native.scene_params(*args, tiles=(ta, tb))):
            mosaic2.array[:, r0:r1, c0:c1] = mine[:, r0:r1, c0:c1]
        mosaic2.publish(1)
        mosaic2.wait(1)
        shared2 = mosaic2.array.copy()
        dist.barrier()
        mosaic2.close()
        if rank == 0:
            q.put((full.numpy(), shared, shared2))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_strip_gather_equals_whole_scene():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, shared, shared2 = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = load_golden('solver_96_t16_s12_ws5')
    d, s = O.image_cut_solver(g['img1'], g['img2'], tuple(g['image_size']), tuple(g['stride']), int(g['ws']),
                              tuple(str(m) for m in g['modes']), bool(g['sub_pix']))
    assert np.array_equal(full[:-1], d, equal_nan=True)
    assert np.array_equal(full[-1], s, equal_nan=True)
    # the mosaic assembled in shared host memory (every rank writes its own strip) is the same
    assert np.array_equal(shared, full, equal_nan=True)
    # ... and so is the one assembled from tile ranges through their owned rectangles
    assert np.array_equal(shared2, full, equal_nan=True)
