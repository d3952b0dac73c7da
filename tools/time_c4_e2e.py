"""Where the end-to-end time of config 4 goes: allocation of the page-locked result arrays,
dm_solve_scene_host itself (pre-allocated arrays), solve_batch, then the post-hoc refinement."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.image_cut_solver import solve_batch, pinned_empty
from deepmatching_stereo_matching_b200.sub_pix_cal import sub_pix_cal_batch
i1, i2 = bench.make_scene('c4')
h1 = pinned_empty(i1.shape, np.uint8); h1[...] = i1
h2 = pinned_empty(i2.shape, np.uint8); h2[...] = i2
kw = dict(image_size=[32, 32], stride=[32, 32], window_size=5, degree_map_mode=bench.MODES, sub_pix=True, devices=[0])
def t(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r
ms, _ = t(lambda: (pinned_empty((64, 2, 448, 448), np.float64), pinned_empty((64, 448, 448), np.float64)))
print('pinned_empty of the result arrays %.2f ms' % ms)
prm = _native.scene_params((512, 512), [32, 32], [32, 32], 5, bench.FEATURE, bench.MODES, True, None, -1, n_scenes=64)
ctx = _native.Context()
dm_, om_ = pinned_empty((64, 2, 448, 448), np.float64), pinned_empty((64, 448, 448), np.float64)
ms, _ = t(lambda: ctx.solve_host(prm, h1, h2, dm_, om_))
print('dm_solve_scene_host into pre-allocated arrays %.2f ms  (DM_STREAM_CHUNKS=%s)' % (ms, os.environ.get('DM_STREAM_CHUNKS', 'default')))
d1, d2 = torch.from_numpy(h1).cuda(), torch.from_numpy(h2).cuda()
dd = torch.zeros((64, 2, 448, 448), dtype=torch.float64, device='cuda'); oo = torch.zeros((64, 448, 448), dtype=torch.float64, device='cuda')
ms, _ = t(lambda: ctx.solve_device(prm, d1, d2, dd, oo))
print('dm_solve_scene (device-resident) %.2f ms' % ms)
ms, (d, s) = t(lambda: solve_batch(h1, h2, **kw))
print('solve_batch %.2f ms' % ms)
ms, r = t(lambda: sub_pix_cal_batch(d, s, [1, 0]), n=10)
print('sub_pix_cal_batch %.2f ms' % ms)
