import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


TILE_CASES = ['tile_16x16_ws5_noise', 'tile_16x16_ws5_sine', 'tile_8x32_ws3', 'tile_32x8_ws5',
              'tile_16x16_ws15', 'tile_8x8_ws5_flat', 'tile_16x16_ws5_ccoeff']
SOLVER_CASES = ['solver_96_t16_s12_ws5', 'solver_80x112_t16_s16_ws3', 'solver_c1_256_t32_s32_ws5']   # the last one: BASELINE.json configs[0]

# tolerance of the float32 ZNCC + min-max against the live reference: OpenCV's own
# float32 cross-correlation is ~2e-6 (up to 4e-5 on low-contrast ws=3 patches) away from
# the exact value (measured again by tests/golden/make_golden.py).
CO_MAP_ATOL = 6e-5
# 3x3 windows of the blurred texture have almost no contrast; there OpenCV's float32
# cross-correlation noise is amplified by 1/(K*sigma^2) and reaches 1.5e-4.
CO_MAP_ATOL_CASE = {'tile_8x32_ws3': 5e-4}


def co_map_atol(name):
    return CO_MAP_ATOL_CASE.get(name, CO_MAP_ATOL)


@pytest.fixture(scope='session')
def has_cuda():
    import torch
    return torch.cuda.is_available()
