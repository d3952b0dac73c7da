"""Determinism stress test of the correlation engines: repeat the same launch and compare bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.synth import stereo_pair
from deepmatching_stereo_matching_b200 import image_cut_solver as ics

lib = _native.lib()
bad = 0
for (shape, T, s, ws) in [((150, 182), 32, 32, 5), ((300, 300), 64, 60, 15), ((200, 168), 32, 30, 5), ((96, 96), 16, 12, 5)]:
    i1, i2 = stereo_pair(shape, seed=5, mode='sine', amp=4)
    for fused in (1, 0):
        ref = None
        nbad = 0
        for rep in range(12):
            sv = ics.ImageCutSolver(i1, i2, image_size=[T, T], stride=[s, s], window_size=ws, degree_map_mode=['elevation', 'elevation2'], sub_pix=True)
            sv.log_flg = False; sv.fused = fused
            d, sc = sv()
            cur = (d.copy(), sc.copy())
            if ref is None: ref = cur
            elif not (np.array_equal(ref[0], cur[0], equal_nan=True) and np.array_equal(ref[1], cur[1], equal_nan=True)):
                nbad += 1
                if nbad == 1:
                    diff = np.argwhere(ref[1] != cur[1])
                    print('   first diffs (score plane):', diff[:6].tolist(), 'count', len(diff))
        print(shape, T, ws, 'fused' if fused else 'materialising', 'nondeterministic runs:', nbad, '/ 11')
        bad += nbad
print('TOTAL', bad)
