"""Pins the numpy oracle against golden vectors produced by the live reference."""
import numpy as np
import pytest

from conftest import load_golden, TILE_CASES, SOLVER_CASES, CO_MAP_ATOL, co_map_atol
from oracle import dm_oracle as O


def _levels(g):
    return [g['level%d' % k] for k in range(int(g['nlevels']))]


@pytest.mark.parametrize('name', TILE_CASES)
def test_co_map_close_to_reference(name):
    g = load_golden(name)
    method = O.FEATURE_NAMES[str(g['feature'])]
    co = O.initial_co_map(g['img1'], g['img2'], int(g['ws']), method)
    ref = g['co_map'].astype(np.float64)
    assert co.shape == ref.shape
    assert np.array_equal(np.isnan(co), np.isnan(ref))
    assert np.nanmax(np.abs(co - ref)) <= co_map_atol(name)


@pytest.mark.parametrize('name', TILE_CASES)
def test_pyramid_bit_exact_given_co_map(name):
    g = load_golden(name)
    lst, it, n = O.pyramid(g['co_map'].astype(np.float64))
    assert it == int(g['iteration']) and n == int(g['N_map'])
    for a, b in zip(lst, _levels(g)):
        assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize('name', TILE_CASES + ['tile_32x32_ws5'])
def test_matching_bit_exact_given_pyramid(name):
    g = load_golden(name)
    if name == 'tile_32x32_ws5':
        pytest.skip('level 0/1 not stored for this fixture')
    lst = _levels(g)
    assert np.array_equal(O.matching(lst, False), g['map_nosub'], equal_nan=True)
    assert np.array_equal(O.matching(lst, True), g['map_sub'], equal_nan=True)
    l32 = [x.astype(np.float32) for x in lst]
    assert np.array_equal(O.matching(l32, False), g['map_f32_nosub'], equal_nan=True)
    assert np.array_equal(O.matching(l32, True), g['map_f32_sub'], equal_nan=True)


@pytest.mark.parametrize('name', TILE_CASES)
def test_cal_map_bit_exact(name):
    g = load_golden(name)
    for m in O.MODES:
        assert np.array_equal(O.cal_map(g['map_sub'], m), g['cal_' + m], equal_nan=True)


def test_end_to_end_32x32():
    g = load_golden('tile_32x32_ws5')
    cm = O.correlation_map(g['img1'], g['img2'], int(g['ws']))
    assert cm['N_map'] == int(g['N_map']) and cm['iteration'] == int(g['iteration'])
    for k in range(2, int(g['nlevels'])):
        assert np.allclose(cm['co_map_list'][k], g['level%d' % k], rtol=1e-3, atol=1e-5)
    out = O.matching(cm['co_map_list'], False)
    bad = np.mean((out[:2] != g['map_nosub'][:2]).any(0))
    assert bad <= 0.001


@pytest.mark.parametrize('name', SOLVER_CASES)
def test_image_cut_solver(name):
    g = load_golden(name)
    d, s = O.image_cut_solver(g['img1'], g['img2'], tuple(g['image_size']), tuple(g['stride']), int(g['ws']),
                              tuple(str(m) for m in g['modes']), bool(g['sub_pix']))
    assert d.shape == g['d_map'].shape and s.shape == g['out_map'].shape
    frac = np.mean(np.abs(d - g['d_map']) > 0.5)
    assert frac <= 0.001
    close = np.abs(d - g['d_map']) <= 1e-3 * np.maximum(1.0, np.abs(g['d_map']))
    assert close.mean() >= 0.999
    assert np.allclose(s, g['out_map'], rtol=1e-3, atol=1e-4) or np.mean(np.abs(s - g['out_map']) > 1e-3) <= 0.001


def test_image_cut_solver_strip_partition():
    g = load_golden('solver_96_t16_s12_ws5')
    args = (g['img1'], g['img2'], tuple(g['image_size']), tuple(g['stride']), int(g['ws']),
            tuple(str(m) for m in g['modes']), bool(g['sub_pix']))
    full_d, full_s = O.image_cut_solver(*args)
    ln = int(g['len'][0])
    merged = np.full_like(full_d, np.nan)
    for lo, hi in [(0, ln // 2), (ln // 2, ln)]:
        d, _ = O.image_cut_solver(*args, tile_rows=(lo, hi))
        m = ~np.isnan(d)
        merged[m] = d[m]                 # higher strip pasted last wins the overlap rows
    assert np.array_equal(merged, full_d, equal_nan=True)


def test_sub_pix_cal_bit_exact():
    g = load_golden('sub_pix_cal')
    for d in (0, 1):
        assert np.array_equal(O.sub_pix_cal(g['arr'], g['co'], direction=d), g['out_dir%d' % d], equal_nan=True)
    assert np.array_equal(O.sub_pix_cal(g['arr'], g['co'], direction=0, ratio=1.), g['out_ratio1'], equal_nan=True)
    assert np.isnan(g['out_dir0']).sum() > 0
    assert np.array_equal(O.image_threshold(g['arr'], [-1, 1]), g['thr'])
    for name in SOLVER_CASES[:1]:
        s = load_golden(name)
        for i, m in enumerate(s['modes']):
            direction = 1 if str(m) == 'elevation' else 0
            assert np.array_equal(O.sub_pix_cal(s['d_map'][i], s['out_map'], direction=direction),
                                  s['spc_' + str(m)], equal_nan=True)


def test_feature_value_general_sizes():
    g = load_golden('feature_value')
    a = O.feature_value(g['patch'], g['img'])
    assert a.dtype == np.float32 and a.shape == g['normed_49'].shape
    assert np.abs(a - g['normed_49']).max() <= CO_MAP_ATOL
    b = O.feature_value(g['patch'], g['img'], O.TM_CCOEFF)
    assert np.abs(b - g['ccoeff_49']).max() <= CO_MAP_ATOL
    c = O.feature_value(g['small'], g['img'])
    assert np.abs(c - g['normed_small']).max() <= CO_MAP_ATOL


def test_against_live_cv2_when_present():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(24, 28), dtype=np.uint8)
    for ws in (3, 5, 15):
        patch = img[2:2 + ws, 5:5 + ws].copy()
        ref = cv2.matchTemplate(patch, img, cv2.TM_CCOEFF_NORMED)
        ref = (ref - ref.min()) / (ref.max() - ref.min())
        assert np.abs(O.feature_value(patch, img) - ref).max() <= CO_MAP_ATOL


def test_raw_read(tmp_path):
    a = (np.arange(12 * 10) % 256).astype(np.uint8).reshape(10, 12)
    p = tmp_path / 'x.raw'
    a.tofile(p)
    assert np.array_equal(O.raw_read(str(p), size=(12, 10)), a)
    assert np.array_equal(O.raw_read(str(p), size=(12, 10), rate=2), (a.astype(np.int8) * 2).astype(np.uint8))


FILTER_CASES = ['filter_16x16_ws5_sine', 'filter_16x16_ws3_unrelated']


@pytest.mark.parametrize('name', FILTER_CASES)
def test_oracle_displacement_filter_vs_reference(name):
    """Matching(filtering=True) of the live reference (misc/Matching.py:224-255) on a float32
    pyramid, 16 settings x sub_pix on/off: the oracle must reproduce every map bit for bit,
    including the numpy index rules of the parabola fit after a filter on level 0."""
    g = load_golden(name)
    lv = [g['level%d' % k] for k in range(int(g['nlevels']))]
    keys = [k for k in g if k.startswith('map_')]
    assert len(keys) == 32 and len(g['crashed']) == 0
    changed = 0
    plain = O.matching(lv, False)
    for k in keys:
        _, mode, num, win, sub = k.split('_')
        got = O.matching(lv, sub == 'sub', filtering=True, filtering_num=int(num[1:]), filter_window_size=int(win[1:]),
                         filtering_mode=mode)
        assert np.array_equal(got, g[k], equal_nan=True), k
        changed += int((g[k][:2].astype(np.int64) != plain[:2].astype(np.int64)).any())
    assert changed > 0          # the fixtures do exercise the filter


def test_oracle_filter_rejects_non_square_maps():
    mp = np.zeros((3, 4, 8))
    with pytest.raises(ValueError):
        O.match_filter(mp, 3, 'median')
    small = np.zeros((3, 2, 8))
    assert np.array_equal(O.match_filter(small, 3, 'median'), small)     # smaller than the window: untouched


def test_oracle_bilateral_filter_vs_cv2():
    """optimize_looper.py:76-77: cv2.bilateralFilter on the uint8 cast of a disparity plane.  The
    oracle restates OpenCV's own algorithm: bit exact against cv2 with IPP off; the IPP build's
    output may differ from that by one grey level (OpenCV against itself)."""
    g = load_golden('bilateral')
    assert np.array_equal(g['plane'].astype('uint8'), g['img'])
    n = 0
    for k in g:
        if not k.startswith('out_'):
            continue
        _, kind, d, sc, ss = k.split('_')
        got = O.bilateral_filter_u8(g['img'], int(d[1:]), float(sc), float(ss))
        if kind == 'plain':
            assert np.array_equal(got, g[k]), k
        else:
            assert np.abs(got.astype(int) - g[k].astype(int)).max() <= 1, k
        n += 1
    assert n == 8


def test_gauss_seidel_wavefront_restatement_is_bit_exact():
    """The reference's sequential in-place Gauss-Seidel loops (misc/optimize_loop.py:15-37,
    misc/opt_loop.py:16-85) restated diagonal by diagonal: same bits as the live reference, including
    the sequentially accumulated error."""
    g = load_golden('gauss_seidel')
    size = g['d'].shape
    for k in range(3):
        alpha, exclusion, loops = g['ol%d_cfg' % k]
        x = g['d'].copy()
        errs = []
        for _ in range(int(loops)):
            x, err = O.optimize_loop(x, g['co'], float(alpha), int(exclusion), size)
            errs.append(err)
        assert np.array_equal(x, g['ol%d_out' % k]) and np.array_equal(np.array(errs), g['ol%d_err' % k])
    for k in range(2):
        s0, s1, exclusion, loops = g['bl%d_cfg' % k]
        gw, cw = O.make_weight(g['d'], int(exclusion), size, np.array([int(s0), int(s1)]))
        assert np.array_equal(gw, g['bl%d_gw' % k]) and np.array_equal(cw, g['bl%d_cw' % k])
        for name, vertical in (('h', False), ('v', True)):
            x = g['d'].copy()
            errs = []
            for _ in range(int(loops)):
                x, err = O.optimize_loop_bilateral(x, cw, gw, g['co'], 0.008, int(exclusion), size, vertical=vertical)
                errs.append(err)
            assert np.array_equal(x, g['bl%d_%s_out' % (k, name)]), (k, name)
            assert np.array_equal(np.array(errs), g['bl%d_%s_err' % (k, name)]), (k, name)
