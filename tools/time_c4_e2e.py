"""Where the end-to-end time of config 4 goes: solve_batch, then the post-hoc refinement."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
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
ms, (d, s) = t(lambda: solve_batch(h1, h2, **kw))
print('solve_batch %.2f ms  (DM_STREAM_CHUNKS=%s)' % (ms, os.environ.get('DM_STREAM_CHUNKS', 'default')))
for ch in (1, 2, 4, 8, 16):
    ms, r = t(lambda: sub_pix_cal_batch(d, s, [1, 0], chunks=ch))
    print('sub_pix_cal_batch chunks=%d %.2f ms' % (ch, ms))
