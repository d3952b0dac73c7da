# -*- coding: utf-8 -*-
"""
Generates tests/golden/*.npz from the LIVE, UNMODIFIED reference (/root/reference).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Inputs come from deepmatching_stereo_matching_b200.synth (seeded).  Every fixture stores
the inputs next to the reference outputs so the tests never need the reference itself.
"""

import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'

# the reference must win the ``misc`` import, the repo only provides synth
sys.path.insert(0, REF)
from misc.Correlation_map import Correlation_map          # noqa: E402
from misc.Matching import Matching                        # noqa: E402
from misc.Calc_difference import Calc_difference          # noqa: E402
from misc.Feature_value import Feature_value              # noqa: E402
from misc.image_cut_solver import ImageCutSolver          # noqa: E402
from misc.sub_pix_cal import sub_pix_cal                  # noqa: E402
from misc.optimize_loop import image_threshold            # noqa: E402
assert sys.modules['misc.Correlation_map'].__file__.startswith(REF)

import importlib.util                                     # noqa: E402
_spec = importlib.util.spec_from_file_location(
    'dm_synth', os.path.join(REPO, 'deepmatching_stereo_matching_b200', 'synth.py'))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)

warnings.filterwarnings('ignore')


def save(name, **kw):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **kw)
    print('%-28s %8.1f KB' % (name, os.path.getsize(path) / 1024))


def tile_case(name, t0, t1, ws, mode, plain, seed, feature='cv2.TM_CCOEFF_NORMED', flat=False,
              keep_levels_from=0):
    e = (ws - 1) // 2
    img1, img2 = synth.stereo_pair((t0 + 2 * e, t1 + 2 * e), seed=seed, mode=mode, amp=3, plain_noise=plain)
    if flat:                                # one flat patch in image 1, one flat window in image 2
        img1 = img1.copy(); img2 = img2.copy()
        img1[2:2 + ws, 3:3 + ws] = 77
        img2[1:1 + ws, 4:4 + ws] = 200
    co = Correlation_map(img1, img2, window_size=ws, feature_name=feature)
    lst = co()
    out = {'img1': img1, 'img2': img2, 'ws': ws, 'feature': feature,
           'N_map': co.N_map, 'iteration': co.iteration, 'nlevels': len(lst)}
    if keep_levels_from == 0:
        out['co_map'] = co.co_map.astype(np.float32)          # exact: float32-valued doubles
        assert np.array_equal(out['co_map'].astype(np.float64), co.co_map, equal_nan=True)
    for k in range(keep_levels_from, len(lst)):
        out['level%d' % k] = lst[k]
    out['map_nosub'] = Matching(co, sub_pix=False)()
    out['map_sub'] = Matching(co, sub_pix=True)()
    for m in ['elevation', 'elevation2', 'distance']:
        out['cal_' + m] = Calc_difference.cal_map(out['map_sub'], m)

    # backtracking on float32 copies of the pyramid (the dtype the GPU pipeline holds)
    class Stub:
        pass
    st = Stub()
    st.co_map_list = [x.astype(np.float32) for x in lst]
    st.N_map = co.N_map
    out['map_f32_nosub'] = Matching(st, sub_pix=False)()
    out['map_f32_sub'] = Matching(st, sub_pix=True)()
    save(name, **out)


def solver_case(name, shape, image_size, stride, ws, modes, sub_pix, seed, mode='sine', amp=3):
    img1, img2 = synth.stereo_pair(shape, seed=seed, mode=mode, amp=amp)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        s = ImageCutSolver(img1, img2, image_size=list(image_size), stride=list(stride), window_size=ws,
                           degree_map_mode=list(modes), sub_pix=sub_pix)
        d_map, out_map = s()
    out = dict(img1=img1, img2=img2, image_size=np.array(image_size), stride=np.array(stride), ws=ws,
               modes=np.array(modes), sub_pix=sub_pix, d_map=d_map, out_map=out_map, len=np.array(s.len))
    # post-hoc sub-pixel (misc/sub_pix_cal.py) on the mosaics, direction rule image_cut_solver.py:137
    for i, m in enumerate(modes):
        direction = 1 if m == 'elevation' else 0
        out['spc_' + m] = sub_pix_cal(d_map[i], out_map, direction=direction)
    save(name, **out)


def feature_case():
    rng = np.random.default_rng(7)
    img = synth.texture((120, 100), seed=11)
    patch = img[30:79, 20:69].copy()                       # 49x49 patch as in for_igarss/cor_map.py:33-35
    small = rng.integers(0, 256, size=(7, 5), dtype=np.uint8)
    out = dict(img=img, patch=patch, small=small)
    out['normed_49'] = Feature_value('cv2.TM_CCOEFF_NORMED')(patch, img)
    out['ccoeff_49'] = Feature_value('cv2.TM_CCOEFF')(patch, img)
    out['normed_small'] = Feature_value('cv2.TM_CCOEFF_NORMED')(small, img)
    save('feature_value', **out)


def subpix2d_case():
    rng = np.random.default_rng(5)
    arr = rng.normal(size=(40, 37)) * 2.5
    co = rng.random((40, 37)) * 2
    co[10:20, 10:20] = 0.7                                 # flat score patch -> 0/0 -> NaN survives
    co[25, 5:9] = [0.1, 0.9, 0.1, 0.5]
    out = dict(arr=arr, co=co)
    for d in (0, 1):
        out['out_dir%d' % d] = sub_pix_cal(arr, co, direction=d)
    out['out_ratio1'] = sub_pix_cal(arr, co, direction=0, ratio=1.)
    out['thr'] = image_threshold(arr, threshold=[-1, 1])
    save('sub_pix_cal', **out)


if __name__ == '__main__':
    tile_case('tile_16x16_ws5_noise', 16, 16, 5, 'shift', True, 1)
    tile_case('tile_16x16_ws5_sine', 16, 16, 5, 'sine', False, 2)
    tile_case('tile_8x32_ws3', 8, 32, 3, 'shift', False, 3)
    tile_case('tile_32x8_ws5', 32, 8, 5, 'sine', False, 4)
    tile_case('tile_16x16_ws15', 16, 16, 15, 'shift', False, 5)
    tile_case('tile_8x8_ws5_flat', 8, 8, 5, 'shift', True, 6, flat=True)
    tile_case('tile_16x16_ws5_ccoeff', 16, 16, 5, 'shift', False, 8, feature='cv2.TM_CCOEFF')
    tile_case('tile_32x32_ws5', 32, 32, 5, 'sine', False, 9, keep_levels_from=2)
    solver_case('solver_96_t16_s12_ws5', (96, 96), (16, 16), (12, 12), 5, ('elevation', 'elevation2'), True, 10)
    solver_case('solver_80x112_t16_s16_ws3', (80, 112), (16, 16), (16, 16), 3, ('distance', 'elevation'), False, 12)
    # BASELINE.json configs[0]: 256 x 256 pair, ws 5, image_size 32, stride 32 -> 36 tiles, 192 x 192
    solver_case('solver_c1_256_t32_s32_ws5', (256, 256), (32, 32), (32, 32), 5, ('elevation', 'elevation2'), True, 0, amp=8)
    feature_case()
    subpix2d_case()
