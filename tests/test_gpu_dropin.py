"""The call sequences of the reference's scripts, replayed through the `misc.*` import
paths they use (the scripts themselves live in the reference checkout, not here)."""
import numpy as np
import pytest

from oracle import dm_oracle as O
from parity_util import NEAR_TIE

pytestmark = pytest.mark.gpu


def _pair(shape, seed, amp=3):
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    return stereo_pair(shape, seed=seed, mode='sine', amp=amp)


def test_deep_dem_mathing_sequence(tmp_path):
    """deep_dem_mathing.py:61-75 -- un-tiled crop, window 5, default Matching, cal_map, np.save."""
    from misc.Correlation_map import Correlation_map
    from misc.Matching import Matching
    from misc.Calc_difference import Calc_difference
    img1, img2 = _pair((36, 68), 41)                 # crop 36x68 -> 32x64 patch grid (non-square)
    co_cls = Correlation_map(img1, img2, window_size=5, feature_name='cv2.TM_CCOEFF_NORMED')
    co_cls()
    cls = Matching(co_cls)
    out = cls()
    d_map = Calc_difference.cal_map(out, mode='elevation')
    assert out.shape == (3, 32, 64) and out.dtype == np.float64 and d_map.shape == (32, 64)
    np.save(tmp_path / 'response.npy', out)
    ref = O.correlation_map(img1, img2, 5)
    ro = O.matching(ref['co_map_list'], True)
    assert co_cls.N_map == ref['N_map'] == 32 and co_cls.iteration == ref['iteration'] == 6
    # <= 0.1 % integer disagreement (2 of 2048 pixels), each one a near-tie of the oracle's own decision
    _, margin = O.matching_margins(ref['co_map_list'])
    bad = np.abs(O.cal_map(ro, 'elevation') - d_map) > 0.5
    print('deep_dem_mathing sequence: %d of %d pixels differ, worst margin %.2e' % (bad.sum(), bad.size, margin[bad].max() if bad.any() else 0))
    assert bad.mean() <= 1e-3 and not (bad & ~(margin < NEAR_TIE)).any()
    assert np.allclose(np.load(tmp_path / 'response.npy'), out)
    # co_map_list behaves like the reference's list of float64 arrays
    lst = co_cls.co_map_list
    assert len(lst) == 6 and lst[0].shape == (32, 64, 32, 64) and lst[-1].shape == (1, 2, 1, 2) and lst[2].dtype == np.float64
    assert [x.shape for x in lst][1] == (16, 32, 16, 32)


def test_bad_matching_sequence():
    """bad_matching.py:60-70 -- private methods + argmax over co_map[i, j, i, :]."""
    from misc.Correlation_map import Correlation_map
    img1, img2 = _pair((20, 36), 42, amp=2)
    co_cls = Correlation_map(img1, img2, window_size=5, feature_name='cv2.TM_CCOEFF_NORMED')
    co_cls._create_atomic_patch()
    co_cls._create_simple_initial_co_map()
    dis = np.zeros((co_cls.co_map.shape[0], co_cls.co_map.shape[1]))
    for i in range(co_cls.co_map.shape[0]):
        for j in range(co_cls.co_map.shape[1]):
            dis[i, j] = j - np.argmax(co_cls.co_map[i, j, i, :])
    # the loop above was served from P own-row slices (one dm_row_argmax launch): the P x P map never came to the host
    assert co_cls.co_map._host_arr is None
    assert np.array_equal(co_cls.co_map.row_argmax(), np.arange(dis.shape[1])[None, :] - dis)
    ref = O.initial_co_map(img1, img2, 5)
    rd = np.array([[j - np.argmax(ref[i, j, i, :]) for j in range(ref.shape[1])] for i in range(ref.shape[0])])
    # the row argmax may only differ where the oracle's two best values of that row are closer than the
    # float32 error of co_map (2e-6 against the exact value, twice)
    rows = np.stack([[ref[i, j, i, :] for j in range(ref.shape[1])] for i in range(ref.shape[0])])
    top2 = np.sort(rows, axis=-1)[..., -2:]
    near = (top2[..., 1] - top2[..., 0]) < 1e-5
    print('bad_matching sequence: %d of %d row maxima differ' % ((dis != rd).sum(), dis.size))
    assert not ((dis != rd) & ~near).any() and np.mean(dis != rd) <= 1e-3
    assert co_cls.atomic_patch.shape == (16, 32, 5, 5) and co_cls.atomic_patch.dtype == np.uint8


def test_ex_deepmatching_rawinput_sequence(tmp_path):
    """ex_deepmatching_rawinput.py:50-80 -- RawRead, crop, ImageCutSolver(ws 15, 64, 60), np.save."""
    from misc.image_cut_solver import ImageCutSolver
    from misc.raw_read import RawRead
    a, b = _pair((400, 420), 43, amp=6)
    a.tofile(tmp_path / 'a.raw'); b.tofile(tmp_path / 'b.raw')
    image1 = RawRead.read(str(tmp_path / 'a.raw'), size=(420, 400), rate=1)
    image2 = RawRead.read(str(tmp_path / 'b.raw'), size=(420, 400), rate=1)
    assert np.array_equal(image1, a)
    image1 = image1[20:20 + 340, 30:30 + 340]
    image2 = image2[20:20 + 340, 30:30 + 340]
    ImageCutSolver.image_save(str(tmp_path / 'here.png'), image1, threshold=[0, 255])
    names = ['elevation', 'elevation2']
    solver = ImageCutSolver(image1, image2, degree_map_mode=names, window_size=15, image_size=[64, 64], stride=[60, 60],
                            sub_pix=False, filtering=False, filtering_window_size=3, filtering_num=4, filtering_mode='median')
    res_list, correlation_map = solver()
    for k, name in enumerate(names):
        np.save(tmp_path / (name + '.npy'), res_list[k])
    np.save(tmp_path / 'correlation.npy', correlation_map)
    assert res_list.shape == (2, 244, 244) and res_list.dtype == np.float64 and correlation_map.shape == (244, 244)
    assert solver.len == [4, 4]
    rd, rs = O.image_cut_solver(np.ascontiguousarray(image1), np.ascontiguousarray(image2), (64, 64), (60, 60), 15, names, False)
    assert np.mean(np.abs(res_list - rd) > 0.5) <= 1e-3
    assert np.mean(np.abs(correlation_map - rs) > 1e-3) <= 1e-3
    # optimize_looper.py:39-45 / sub_pix_cal.__main__ read these files back as 2-D float64 arrays
    assert np.load(tmp_path / 'elevation.npy').dtype == np.float64
    from misc.sub_pix_cal import sub_pix_cal
    sp = sub_pix_cal(np.load(tmp_path / 'elevation.npy'), np.load(tmp_path / 'correlation.npy'), direction=1)
    assert np.array_equal(sp, O.sub_pix_cal(res_list[0], correlation_map, direction=1), equal_nan=True)


def test_matching_with_filtering_runs():
    """ex_deepmatching_rawinput.py:32-35 flags: the displacement filter (default off) works on
    square grids and changes nothing when the field is smooth."""
    from misc.Correlation_map import Correlation_map
    from misc.Matching import Matching
    img1, img2 = _pair((36, 36), 44, amp=0)
    co = Correlation_map(img1, img2, window_size=5)
    co()
    plain = Matching(co, sub_pix=False)()
    filt = Matching(co, sub_pix=False, filtering=True, filtering_num=3, filtering_mode='median')()
    assert filt.shape == plain.shape
    assert np.mean((filt[:2] != plain[:2]).any(0)) < 0.2
