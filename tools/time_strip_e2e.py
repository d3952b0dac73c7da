"""One rank's share of the C3 scene at N = 8 (545 tiles), on one GPU: device-resident solve against
the host-buffer solve (upload of the rows, solve, finished pixels streamed into page-locked arrays).
DM_STREAM_CHUNKS is read once per process: run once per setting."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.image_cut_solver import pinned_empty

name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
c = bench.CONFIGS[name]
i1, i2 = bench.make_scene(name)
h1 = pinned_empty(i1.shape, np.uint8); h1[...] = i1
h2 = pinned_empty(i2.shape, np.uint8); h2[...] = i2
len0, len1, out = bench.geometry(c)
parts = _native.partition_tile_rows(len0 * len1, world)
a, b = parts[world // 2]
prm = _native.scene_params(c['shape'], [c['T']] * 2, [c['stride']] * 2, c['ws'], bench.FEATURE, bench.MODES, True, tiles=(a, b))
ctx = _native.Context()
dmap = pinned_empty((2, out[0], out[1]), np.float64); omap = pinned_empty(out, np.float64)
d1, d2 = torch.from_numpy(h1).cuda(), torch.from_numpy(h2).cuda()
pl = torch.zeros((3, out[0], out[1]), dtype=torch.float64, device='cuda')
for _ in range(3):
    ctx.solve_device(prm, d1, d2, pl[:-1], pl[-1])
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    ctx.solve_device(prm, d1, d2, pl[:-1], pl[-1])
torch.cuda.synchronize()
dev = (time.perf_counter() - t) / 20
for _ in range(3):
    info = ctx.solve_host(prm, h1, h2, dmap, omap)
t = time.perf_counter()
for _ in range(20):
    info = ctx.solve_host(prm, h1, h2, dmap, omap)
host = (time.perf_counter() - t) / 20
print('%s share %d of %d (tiles %d..%d): device %.3f ms, host e2e %.3f ms, chunk_tiles %d, DM_STREAM_CHUNKS=%s' % (
    name, world // 2, world, a, b, dev * 1e3, host * 1e3, info.chunk_tiles, os.environ.get('DM_STREAM_CHUNKS', 'default')))
