# -*- coding: utf-8 -*-
"""
Golden vectors for Matching._filter (misc/Matching.py:224-255) from the LIVE, UNMODIFIED
reference (/root/reference); build container only:

    python tests/golden/make_golden_filter.py

The displacement filter is off by default (ex_deepmatching_rawinput.py:32-35 exposes it).  It
is only well defined on square patch grids (its d_map is sized (shape[1], shape[1])) and it can
push p_dot out of range, after which the reference raises IndexError -- combinations that
crash the reference are recorded as such and are not part of the parity contract.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REF)
from misc.Correlation_map import Correlation_map          # noqa: E402
from misc.Matching import Matching                        # noqa: E402
assert sys.modules['misc.Correlation_map'].__file__.startswith(REF)

import importlib.util                                     # noqa: E402
_spec = importlib.util.spec_from_file_location(
    'dm_synth', os.path.join(REPO, 'deepmatching_stereo_matching_b200', 'synth.py'))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)
warnings.filterwarnings('ignore')

COMBOS = [(mode, num, win) for mode in ('median', 'average') for num in (3, 4, 5, 9) for win in (3, 5)]


def case(name, t, ws, seed, independent):
    e = (ws - 1) // 2
    if independent:        # two unrelated textures: the matches are outlier-ridden, the filter has work to do
        img1 = synth.texture((t + 2 * e, t + 2 * e), seed=seed)
        img2 = synth.texture((t + 2 * e, t + 2 * e), seed=seed + 1000)
    else:
        img1, img2 = synth.stereo_pair((t + 2 * e, t + 2 * e), seed=seed, mode='sine', amp=3)
    co = Correlation_map(img1, img2, window_size=ws)
    lst = co()

    class Stub:
        pass
    st = Stub()
    st.co_map_list = [x.astype(np.float32) for x in lst]
    st.N_map = co.N_map
    out = {'img1': img1, 'img2': img2, 'ws': ws, 'nlevels': len(lst)}
    for k, x in enumerate(st.co_map_list):
        out['level%d' % k] = x                      # the float32 pyramid both sides backtrack on
    crashed = []
    for mode, num, win in COMBOS:
        key = '%s_n%d_w%d' % (mode, num, win)
        for sub in (False, True):
            try:
                mp = Matching(st, filter_window_size=win, filtering=True, filtering_num=num, filtering_mode=mode, sub_pix=sub)()
                out['map_%s_%s' % (key, 'sub' if sub else 'nosub')] = mp
            except Exception as ex:        # noqa: BLE001 -- the reference's own failure mode
                crashed.append('%s_%s:%s' % (key, 'sub' if sub else 'nosub', type(ex).__name__))
    out['crashed'] = np.array(crashed)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-28s %8.1f KB  crashed: %s' % (name, os.path.getsize(path) / 1024, crashed))


if __name__ == '__main__':
    case('filter_16x16_ws5_sine', 16, 5, 21, False)
    case('filter_16x16_ws3_unrelated', 16, 3, 22, True)
