"""Where does the end-to-end time go?  C call with pinned buffers vs the full ImageCutSolver call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.image_cut_solver import ImageCutSolver, pinned_empty
from deepmatching_stereo_matching_b200.synth import stereo_pair

i1, i2 = stereo_pair((1024, 1024), seed=1, mode='sine', amp=16)
h1 = pinned_empty(i1.shape, np.uint8); h1[...] = i1
h2 = pinned_empty(i2.shape, np.uint8); h2[...] = i2
modes = ['elevation', 'elevation2']
prm = _native.scene_params((1024, 1024), (64, 64), (60, 60), 15, 'cv2.TM_CCOEFF_NORMED', modes, True, None, -1)
ctx = _native.Context()
d = pinned_empty((2, 904, 904), np.float64); s = pinned_empty((904, 904), np.float64)

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - a) / n * 1e3

print('dm_solve_scene_host (pinned in/out)      %.3f ms' % t(lambda: ctx.solve_host(prm, h1, h2, d, s)))
dp = np.empty((2, 904, 904)); sp = np.empty((904, 904))
print('dm_solve_scene_host (pageable out)       %.3f ms' % t(lambda: ctx.solve_host(prm, h1, h2, dp, sp)))
def full():
    sv = ImageCutSolver(h1, h2, image_size=[64, 64], stride=[60, 60], window_size=15, degree_map_mode=modes, sub_pix=True)
    sv.log_flg = False
    return sv()
print('ImageCutSolver(...)()                    %.3f ms' % t(full))
def parts():
    sv = ImageCutSolver(h1, h2, image_size=[64, 64], stride=[60, 60], window_size=15, degree_map_mode=modes, sub_pix=True)
    sv._cut_and_pool()
print('  constructor + _cut_and_pool            %.3f ms' % t(parts))
print('  pinned_empty x2                        %.3f ms' % t(lambda: (pinned_empty((2, 904, 904), np.float64), pinned_empty((904, 904), np.float64))))
dd = torch.empty((3, 904, 904), dtype=torch.float64, device='cuda'); hh = torch.empty((3, 904, 904), dtype=torch.float64, pin_memory=True)
print('  D2H 19.6 MB pinned                     %.3f ms' % t(lambda: hh.copy_(dd, non_blocking=True)))
