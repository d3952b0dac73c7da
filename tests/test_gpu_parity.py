"""Parity of the CUDA path (through the C-ABI / the reference-shaped classes) against the
golden vectors of the live reference and against the numpy oracle.  Needs a B200."""
import numpy as np
import pytest

from conftest import load_golden, TILE_CASES, SOLVER_CASES, co_map_atol
from oracle import dm_oracle as O
from parity_util import scene_report, assert_parity, NEAR_TIE

pytestmark = pytest.mark.gpu

# float32 pyramid on the GPU vs the reference's float64 pyramid (north star: 1e-3 relative)
LEVEL_RTOL, LEVEL_ATOL = 1e-3, 2e-5
# integer-disparity disagreement allowed end to end (north star: <= 0.1 %)
MAX_INDEX_DISAGREEMENT = 1e-3


@pytest.fixture(scope='module')
def dm():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import deepmatching_stereo_matching_b200 as pkg
    from deepmatching_stereo_matching_b200 import _native
    assert _native.lib().dm_device_cc() >= 100
    return pkg


def _levels(g):
    return [g['level%d' % k] for k in range(int(g['nlevels']))]


class Stub(object):
    pass


@pytest.fixture(params=[-1, 0], ids=['cta_pair', 'single_cta'])
def pair_mode(request, dm):
    """Runs a test on both builds of the tcgen05 correlation kernel: CTA pairs
    (tcgen05.mma.cta_group::2, the default where a tile has an even number of work items)
    and one CTA per work item."""
    from deepmatching_stereo_matching_b200 import _native
    _native.check(_native.lib().dm_correlation_set_pair_mode(request.param))
    yield request.param
    _native.check(_native.lib().dm_correlation_set_pair_mode(-1))


@pytest.mark.parametrize('engine', [1, 0])
@pytest.mark.parametrize('name', TILE_CASES)
def test_correlation_and_pyramid_vs_reference(dm, name, engine):
    g = load_golden(name)
    co = dm.Correlation_map(g['img1'], g['img2'], window_size=int(g['ws']), feature_name=str(g['feature']))
    co._create_atomic_patch()
    assert np.array_equal(co.atomic_patch, O.atomic_patches(g['img1'], int(g['ws'])))
    co._create_simple_initial_co_map(engine=engine)
    ref = g['co_map'].astype(np.float64)
    got = co.co_map
    assert got.shape == ref.shape and got.dtype == np.float64
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.nanmax(np.abs(got - ref)) <= co_map_atol(name)
    # tighter: against the oracle's exact-integer ZNCC
    exact = O.initial_co_map(g['img1'], g['img2'], int(g['ws']), O.FEATURE_NAMES[str(g['feature'])])
    assert np.nanmax(np.abs(got - exact)) <= 2e-6
    co._multi_level_correlation_pyramid()
    assert co.N_map == int(g['N_map']) and co.iteration == int(g['iteration'])
    assert len(co.co_map_list) == int(g['nlevels'])
    # the reference's own co_map carries OpenCV's float32 cross-correlation noise
    # (co_map_atol); the pyramid is 1-Lipschitz-ish in it, so allow 10x that against the
    # golden levels and hold the tight bound against the oracle's exact pyramid
    exact_levels, _, _ = O.pyramid(exact)
    for a, b, c in zip(co.co_map_list, _levels(g), exact_levels):
        assert a.shape == b.shape
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.allclose(a, b, rtol=LEVEL_RTOL, atol=max(LEVEL_ATOL, 10 * co_map_atol(name)), equal_nan=True)
        assert np.allclose(a, c, rtol=1e-4, atol=5e-6, equal_nan=True)


@pytest.mark.parametrize('t0,t1,ws,n', [(16, 16, 5, 3), (8, 32, 3, 2), (32, 32, 5, 5), (32, 32, 15, 2), (64, 64, 15, 3), (32, 64, 7, 2)])
def test_tcgen05_correlation_agrees_with_simt(dm, t0, t1, ws, n, pair_mode):
    """Both engines accumulate the same exact integer products.  The CUDA-core engine applies the
    -S1' S2'/K term with one FMA; the tensor-core engine gets it from the three correction entries of
    the descriptor rows, whose accumulation rounds by at most a few ulp of the accumulator
    (|acc| < 2^24): raw ZNCC within 5e-7, the un-normalised numerator within 4."""
    import torch
    from deepmatching_stereo_matching_b200 import _native
    from deepmatching_stereo_matching_b200.synth import texture
    lib = _native.lib()
    e2 = ws - 1
    H, W = t0 + e2 + 7, (t1 + e2) + 13 * (n - 1)
    s1 = torch.from_numpy(texture((H, W), seed=31)).cuda()
    s2 = torch.from_numpy(texture((H, W), seed=32, plain_noise=True)).cuda()
    origin = torch.tensor([[k % 7, 13 * k] for k in range(n)], dtype=torch.int32, device='cuda')
    P, kpad = t0 * t1, lib.dm_kpad(ws)
    bufs = []
    for side, sc in ((1, s1), (2, s2)):
        desc = torch.empty((n * P, kpad), dtype=torch.bfloat16, device='cuda')
        stat = torch.empty((n * P * 6,), dtype=torch.float32, device='cuda')
        _native.check(lib.dm_descriptors(_native.ptr(sc), H, W, W, _native.ptr(origin), n, t0, t1, ws, side,
                                         _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
        bufs += [desc, stat]
    for method in (_native.TM_CCOEFF_NORMED, _native.TM_CCOEFF):
        out = []
        for engine in (_native.CORR_SIMT, _native.CORR_UMMA):
            raw = torch.full((n, P, P), float('nan'), dtype=torch.float32, device='cuda')
            _native.check(lib.dm_correlation(_native.ptr(bufs[0]), _native.ptr(bufs[1]), _native.ptr(bufs[2]), _native.ptr(bufs[3]),
                                             n, P, kpad, ws, method, engine, _native.ptr(raw), _native.stream_ptr()))
            torch.cuda.synchronize()
            out.append(raw.cpu().numpy())
        assert not np.isnan(out[1]).any()
        if method == _native.TM_CCOEFF_NORMED:
            assert np.abs(out[0] - out[1]).max() <= 5e-7
        else:
            assert np.abs(out[0] - out[1]).max() <= 4.0 and np.allclose(out[0], out[1], rtol=1e-6, atol=4.0)
        print('ws %d method %d: engines differ in %.4f of the entries, max %.2e' % (ws, method, np.mean(out[0] != out[1]), np.abs(out[0] - out[1]).max()))
    # and the SIMT engine against the oracle for the first tile
    i1 = s1[:t0 + e2, :t1 + e2].cpu().numpy()
    i2 = s2[:t0 + e2, :t1 + e2].cpu().numpy()
    ref = O.match_template_matrix(i1, i2, ws)
    raw = torch.empty((n, P, P), dtype=torch.float32, device='cuda')
    _native.check(lib.dm_correlation(_native.ptr(bufs[0]), _native.ptr(bufs[1]), _native.ptr(bufs[2]), _native.ptr(bufs[3]),
                                     n, P, kpad, ws, _native.TM_CCOEFF_NORMED, _native.CORR_AUTO, _native.ptr(raw), _native.stream_ptr()))
    assert np.abs(raw[0].cpu().numpy() - ref).max() <= 1e-6


@pytest.mark.parametrize('T,ws,n', [(32, 7, 6), (64, 5, 7), (32, 5, 20), (64, 15, 3)])
def test_correlation_engines_are_repeatable(dm, T, ws, n, pair_mode):
    """Race detector for the warp-specialised tcgen05 kernel: the same launch repeated 25 times
    must produce identical bits (few work items per SM and a short K make the MMA, TMA and
    epilogue pipelines run at very different speeds -- this caught an early TMEM-stage release)."""
    import torch
    from deepmatching_stereo_matching_b200 import _native
    from deepmatching_stereo_matching_b200.synth import texture
    lib = _native.lib()
    e2 = ws - 1
    H, W = T + e2 + 9, (T + e2) + 11 * (n - 1)
    s1 = torch.from_numpy(texture((H, W), seed=31)).cuda()
    s2 = torch.from_numpy(texture((H, W), seed=32, plain_noise=True)).cuda()
    origin = torch.tensor([[k % 9, 11 * k] for k in range(n)], dtype=torch.int32, device='cuda')
    P, kpad = T * T, lib.dm_kpad(ws)
    bufs = []
    for side, sc in ((1, s1), (2, s2)):
        desc = torch.empty((n * P, kpad), dtype=torch.bfloat16, device='cuda')
        stat = torch.empty((n * P * 6,), dtype=torch.float32, device='cuda')
        _native.check(lib.dm_descriptors(_native.ptr(sc), H, W, W, _native.ptr(origin), n, T, T, ws, side, _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
        bufs += [desc, stat]
    ref = torch.empty((n * P * P,), dtype=torch.float32, device='cuda')
    _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, _native.CORR_SIMT, _native.ptr(ref), _native.stream_ptr()))
    # 4 = pooled epilogue (test aid)
    psize = n * P * (P // 4) + 8 * n * P
    for engine, size in ((_native.CORR_UMMA, n * P * P), (4, psize)):
        first = None
        for _ in range(25):
            out = torch.full((size,), float('nan'), dtype=torch.float32, device='cuda')
            _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, engine, _native.ptr(out), _native.stream_ptr()))
            torch.cuda.synchronize()
            cur = out.view(torch.int32)
            if first is None:
                first = cur.clone()
                assert not torch.isnan(out).any()
                if engine == _native.CORR_UMMA:
                    assert (out - ref).abs().max().item() <= 5e-7
            else:
                assert torch.equal(cur, first)


@pytest.mark.parametrize('T,ws,n', [(16, 5, 3), (32, 5, 5), (32, 15, 2), (64, 15, 3), (64, 3, 2), (128, 5, 1)])
def test_pooled_epilogue_equals_maxpool_of_raw_zncc(dm, T, ws, n, pair_mode):
    """The pooled tcgen05 epilogue against torch's max_pool2d(3, 2, 1) of the raw ZNCC the same
    tensor-core engine materialises (identical accumulators): clamp and the row factor are
    monotone, so the pooled map, the per-patch minimum and the per-patch maximum of the pooled map
    must agree bit for bit; and within 5e-7 of the CUDA-core engine's."""
    import torch
    from deepmatching_stereo_matching_b200 import _native
    from deepmatching_stereo_matching_b200.synth import texture
    lib = _native.lib()
    e2 = ws - 1
    H, W = T + e2 + 9, (T + e2) + 11 * (n - 1)
    s1 = torch.from_numpy(texture((H, W), seed=41)).cuda()
    s2 = torch.from_numpy(texture((H, W), seed=42, plain_noise=True)).cuda()
    origin = torch.tensor([[k % 9, 11 * k] for k in range(n)], dtype=torch.int32, device='cuda')
    P, kpad = T * T, lib.dm_kpad(ws)
    bufs = []
    for side, sc in ((1, s1), (2, s2)):
        desc = torch.empty((n * P, kpad), dtype=torch.bfloat16, device='cuda')
        stat = torch.empty((n * P * 6,), dtype=torch.float32, device='cuda')
        _native.check(lib.dm_descriptors(_native.ptr(sc), H, W, W, _native.ptr(origin), n, T, T, ws, side, _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
        bufs += [desc, stat]
    for method in (_native.TM_CCOEFF_NORMED, _native.TM_CCOEFF):
        raw = torch.empty((n * P, 1, T, T), dtype=torch.float32, device='cuda')
        simt = torch.empty((n * P, 1, T, T), dtype=torch.float32, device='cuda')
        _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, method, _native.CORR_SIMT, _native.ptr(simt), _native.stream_ptr()))
        _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, method, _native.CORR_UMMA, _native.ptr(raw), _native.stream_ptr()))
        if method == _native.TM_CCOEFF_NORMED:
            assert (raw - simt).abs().max().item() <= 5e-7
        want = torch.nn.functional.max_pool2d(raw, 3, 2, 1).reshape(n * P, P // 4)
        want_min = raw.reshape(n * P, P).min(dim=1).values
        want_max = want.max(dim=1).values
        for engine in (4,):
            out = torch.full((n * P * (P // 4) + 8 * n * P,), float('nan'), dtype=torch.float32, device='cuda')
            _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, method, engine, _native.ptr(out), _native.stream_ptr()))
            torch.cuda.synchronize()
            pooled = out[:n * P * (P // 4)].reshape(n * P, P // 4)
            rmin = out[n * P * (P // 4):n * P * (P // 4) + 4 * n * P].reshape(n * P, 4).min(dim=1).values
            rmax = out[n * P * (P // 4) + 4 * n * P:].reshape(n * P, 4).max(dim=1).values
            assert not torch.isnan(out).any()
            assert torch.equal(pooled, want), 'engine %d method %d: %d of %d pooled values differ' % (engine, method, int((pooled != want).sum()), pooled.numel())
            assert torch.equal(rmin, want_min)
            assert torch.equal(rmax, want_max)


@pytest.mark.parametrize('name', TILE_CASES)
def test_aggregate_kernel_given_reference_level(dm, name):
    import torch
    from deepmatching_stereo_matching_b200 import _native
    g = load_golden(name)
    lv = _levels(g)
    for k in range(len(lv) - 1):
        src = torch.from_numpy(lv[k].astype(np.float32)).cuda()
        a, b, c, d = src.shape
        out = torch.empty((a // 2, b // 2, c // 2, d // 2), dtype=torch.float32, device='cuda')
        _native.check(_native.lib().dm_aggregate(_native.ptr(src), 1, a, b, c, d, 1, _native.ptr(out), _native.stream_ptr()))
        got = out.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(lv[k + 1]))
        assert np.allclose(got, lv[k + 1], rtol=2e-5, atol=1e-6, equal_nan=True)
        # un-rectified variant == oracle.aggregate
        _native.check(_native.lib().dm_aggregate(_native.ptr(src), 1, a, b, c, d, 0, _native.ptr(out), _native.stream_ptr()))
        ref = O.aggregate(lv[k].astype(np.float32).astype(np.float64))
        assert np.allclose(out.cpu().numpy(), ref, rtol=1e-6, atol=1e-7, equal_nan=True)


def test_aggregate_batched_and_banded(dm):
    """n > 1, wide rows (banded staging) and a non-square grid against the oracle."""
    import torch
    from deepmatching_stereo_matching_b200 import _native
    rng = np.random.default_rng(0)
    for (n, a, b, c, d) in [(3, 4, 4, 64, 64), (2, 2, 2, 128, 128), (2, 2, 4, 8, 32), (5, 8, 8, 8, 8), (1, 2, 2, 2, 2), (2, 2, 2, 6, 6)]:
        x = rng.random((n, a, b, c, d)).astype(np.float32)
        src = torch.from_numpy(x).cuda()
        out = torch.empty((n, a // 2, b // 2, c // 2, d // 2), dtype=torch.float32, device='cuda')
        _native.check(_native.lib().dm_aggregate(_native.ptr(src), n, a, b, c, d, 1, _native.ptr(out), _native.stream_ptr()))
        ref = np.stack([O.rectify(O.aggregate(x[i].astype(np.float64))) for i in range(n)])
        assert np.allclose(out.cpu().numpy(), ref, rtol=2e-5, atol=1e-6), (n, a, b, c, d)


@pytest.mark.parametrize('name', TILE_CASES)
def test_backtracking_bit_exact(dm, name):
    g = load_golden(name)
    lv = _levels(g)
    st = Stub()
    st.co_map_list = lv                      # float64, as the reference holds it
    st.N_map = int(g['N_map'])
    assert np.array_equal(dm.Matching(st, sub_pix=False)(), g['map_nosub'], equal_nan=True)
    got = dm.Matching(st, sub_pix=True)()
    assert got.dtype == np.float64 and got.shape == g['map_sub'].shape
    assert np.array_equal(got[2], g['map_sub'][2], equal_nan=True)
    assert np.array_equal(got[:2].astype(np.int64), g['map_sub'][:2].astype(np.int64))
    assert np.allclose(got, g['map_sub'], rtol=0, atol=1e-12, equal_nan=True)
    st.co_map_list = [x.astype(np.float32) for x in lv]
    assert np.array_equal(dm.Matching(st, sub_pix=False)(), g['map_f32_nosub'], equal_nan=True)
    assert np.array_equal(dm.Matching(st, sub_pix=True)(), g['map_f32_sub'], equal_nan=True)


@pytest.mark.parametrize('name', ['filter_16x16_ws5_sine', 'filter_16x16_ws3_unrelated'])
def test_displacement_filter_bit_exact(dm, name):
    """Matching(filtering=True): dm_match_filter between the levels (misc/Matching.py:224-255)
    against the live reference's maps on the same float32 pyramid -- all 16 settings, with
    and without the parabola fit (which sees matches one step outside the map after a
    level-0 filter)."""
    g = load_golden(name)
    st = Stub()
    st.co_map_list = [g['level%d' % k] for k in range(int(g['nlevels']))]
    st.N_map = st.co_map_list[0].shape[0]
    for k in [k for k in g if k.startswith('map_')]:
        _, mode, num, win, sub = k.split('_')
        got = dm.Matching(st, filter_window_size=int(win[1:]), filtering=True, filtering_num=int(num[1:]),
                          filtering_mode=mode, sub_pix=(sub == 'sub'))()
        assert np.array_equal(got, g[k], equal_nan=True), k


@pytest.mark.parametrize('mode', ['median', 'average'])
@pytest.mark.parametrize('h,win', [(32, 3), (32, 9), (32, 11), (32, 15), (16, 13), (64, 31)])
def test_match_filter_kernel_any_window(dm, mode, h, win):
    """dm_match_filter on random displacement fields against the oracle, including the windows
    above 9 x 9 that the kernel ranks by bisection (no host path for any window size)."""
    import torch
    from deepmatching_stereo_matching_b200 import _native
    rng = np.random.default_rng(7 * h + win)
    n = 3
    ii, jj = np.meshgrid(np.arange(h), np.arange(h), indexing='ij')
    match = np.stack([np.stack([ii + rng.integers(-6, 7, (h, h)), jj + rng.integers(-6, 7, (h, h))]) for _ in range(n)]).astype(np.int32)
    dev = torch.from_numpy(match).cuda()
    out = torch.empty_like(dev)
    _native.check(_native.lib().dm_match_filter(_native.ptr(dev), n, h, h, win, _native.FILTER_IDS[mode], _native.ptr(out), _native.stream_ptr()))
    got = out.cpu().numpy()
    for k in range(n):
        mp = np.concatenate([match[k].astype(np.float64), np.zeros((1, h, h))], 0)
        want = O.match_filter(mp, win, mode)[:2]
        assert np.array_equal(got[k], want.astype(np.int32)), (mode, h, win, k)


def test_filter_on_non_square_maps_raises_like_the_reference(dm):
    """misc/Matching.py:235-248 fails on non-square maps (broadcast error or round(NaN)); the
    class raises ValueError as well instead of falling back to a host computation."""
    rng = np.random.default_rng(5)
    st = Stub()
    st.co_map_list = [rng.random((8 >> k, 32 >> k, 8 >> k, 32 >> k)).astype(np.float32) for k in range(4)]
    st.N_map = 8
    with pytest.raises(ValueError):
        dm.Matching(st, filter_window_size=3, filtering=True, filtering_num=4, filtering_mode='median', sub_pix=False)()
    img1 = (rng.random((60, 120)) * 255).astype(np.uint8)
    with pytest.raises(ValueError):
        s = dm.ImageCutSolver(img1, img1.copy(), image_size=[8, 32], stride=[8, 32], window_size=3, filtering=True,
                              filtering_window_size=3, filtering_num=4, filtering_mode='median')
        s.log_flg = False
        s()


@pytest.mark.parametrize('mode,num,win,fused_expected', [('median', 3, 3, 1), ('average', 4, 3, 1), ('median', 9, 5, 0), ('average', 2, 9, 1)])
def test_image_cut_solver_with_displacement_filter(dm, mode, num, win, fused_expected):
    """ex_deepmatching_rawinput.py:32-35 flags through ImageCutSolver: the batched solver runs the
    filter kernel between the levels (fused path while the filter stays above level 0, else
    the materialising path) and must equal the tile-by-tile class path; both against the oracle."""
    from deepmatching_stereo_matching_b200.synth import texture
    img1 = texture((120, 120), seed=61)
    img2 = np.roll(texture((120, 120), seed=61), 2, axis=1)
    img2[40:80, 30:90] = texture((40, 60), seed=62)          # a block of outliers for the filter to work on
    kw = dict(image_size=[16, 16], stride=[14, 14], window_size=5, degree_map_mode=['elevation', 'elevation2'], sub_pix=True,
              filtering=True, filtering_window_size=win, filtering_num=num, filtering_mode=mode)
    s = dm.ImageCutSolver(img1, img2, **kw)
    s.log_flg = False
    d, sc = s()
    assert s.info.used_fused == fused_expected
    s2 = dm.ImageCutSolver(img1, img2, **kw)
    s2.log_flg = False
    s2._cut_and_pool()
    s2._execute_matching_per_tile(list(d.shape[1:]))
    _same_up_to_accumulation_rounding(s2.d_map, s2.out_map, d, sc, exact=not fused_expected)
    # "bit-exact given identical correlation inputs": the oracle's top-down pass + filter on the GPU's own
    # (float32) pyramid of a tile must reproduce the GPU's matches exactly -- whatever differs from the
    # float64 oracle below comes from the float32 level values alone
    for (ty, tx) in ((0, 0), (42, 28), (56, 84)):           # the last two lie in / at the outlier block
        co = dm.Correlation_map(img1[ty:ty + 20, tx:tx + 20], img2[ty:ty + 20, tx:tx + 20], window_size=5)
        lv32 = [np.asarray(x, dtype=np.float32) for x in co()]
        got = dm.Matching(co, sub_pix=False, filtering=True, filter_window_size=win, filtering_num=num, filtering_mode=mode)()
        want = O.matching(lv32, False, filtering=True, filtering_num=num, filter_window_size=win, filtering_mode=mode)
        assert np.array_equal(got, want, equal_nan=True)
    rd, rs = O.image_cut_solver(img1, img2, (16, 16), (14, 14), 5, ('elevation', 'elevation2'), True, filtering=(num, win, mode))
    bad = np.abs(d - rd) > 0.5
    inside = np.zeros(d.shape[1:], bool)
    inside[40 - 16:80 + 4, 30 - 16:90 + 4] = True           # pixels whose tile or filter window can see the uncorrelated block
    print('filter %s num %d win %d: disagreement %.5f outside the outlier block, %.5f in its neighbourhood' % (
        mode, num, win, bad[:, ~inside].mean(), bad[:, inside].mean()))
    assert bad[:, ~inside].mean() <= MAX_INDEX_DISAGREEMENT
    assert bad.mean() <= 5e-3                               # uncorrelated texture: the matches are near-ties by construction
    plain, _ = dm.ImageCutSolver(img1, img2, **dict(kw, filtering=False))()
    if win <= 16 >> (5 - num if num < 5 else 0):             # some filtered map is at least as large as the window
        assert np.mean(np.abs(plain - d) > 0.5) > 0          # ... and the filter did change the field
    else:
        assert np.array_equal(plain, d, equal_nan=True)


@pytest.mark.parametrize('name', TILE_CASES)
def test_cal_map_bit_exact(dm, name):
    g = load_golden(name)
    for m in O.MODES:
        assert np.array_equal(dm.Calc_difference.cal_map(g['map_sub'], m), g['cal_' + m], equal_nan=True)


def test_sub_pix_cal_bit_exact(dm):
    g = load_golden('sub_pix_cal')
    for d in (0, 1):
        assert np.array_equal(dm.sub_pix_cal(g['arr'], g['co'], direction=d), g['out_dir%d' % d], equal_nan=True)
    assert np.array_equal(dm.sub_pix_cal(g['arr'], g['co'], direction=0, ratio=1.), g['out_ratio1'], equal_nan=True)
    s = load_golden(SOLVER_CASES[0])
    for i, m in enumerate(s['modes']):
        direction = 1 if str(m) == 'elevation' else 0
        assert np.array_equal(dm.sub_pix_cal(s['d_map'][i], s['out_map'], direction=direction), s['spc_' + str(m)], equal_nan=True)


def test_feature_value_general_sizes(dm):
    g = load_golden('feature_value')
    a = dm.Feature_value('cv2.TM_CCOEFF_NORMED')(g['patch'], g['img'])
    assert a.dtype == np.float32 and a.shape == g['normed_49'].shape
    assert np.abs(a - g['normed_49']).max() <= 6e-5
    assert np.abs(a - O.feature_value(g['patch'], g['img'])).max() <= 1e-6
    b = dm.Feature_value('cv2.TM_CCOEFF')(g['patch'], g['img'])
    assert np.abs(b - g['ccoeff_49']).max() <= 6e-5
    c = dm.Feature_value()(g['img'], g['small'])           # swapped argument order, like cv2
    assert np.abs(c - g['normed_small']).max() <= 6e-5


def _same_up_to_accumulation_rounding(d_a, s_a, d_b, s_b, exact):
    """The batched solver against the tile-by-tile class path.  On the materialising path both run the same
    kernels on the same values: bit for bit.  The fused path recomputes the level-0 values of its last
    step from exact integer dot products, while the class path reads them from the map the tensor-core
    kernel materialised (correction terms accumulated with a few ulp of rounding): same matches, scores
    within 1e-6, sub-pixel shifts within 1e-4 except where the parabola is ill conditioned."""
    assert np.array_equal(np.isnan(d_a), np.isnan(d_b)) and np.array_equal(np.isnan(s_a), np.isnan(s_b))
    if exact:
        assert np.array_equal(d_a, d_b, equal_nan=True) and np.array_equal(s_a, s_b, equal_nan=True)
        return
    ok = ~np.isnan(d_a)
    print('batched vs tile by tile: %.5f of the values differ by more than 0.5, %.5f by more than 1e-4' % (
        np.mean(np.abs(d_a - d_b)[ok] > 0.5), np.mean(np.abs(d_a - d_b)[ok] > 1e-4)))
    assert np.mean(np.abs(d_a - d_b)[ok] > 0.5) <= 1e-3
    assert np.mean(np.abs(d_a - d_b)[ok] > 1e-4) <= 2e-3
    so = ~np.isnan(s_a)
    assert np.mean(np.abs(s_a - s_b)[so] > 1e-6) <= 1e-3


def _end_to_end_tile(dm, img1, img2, ws, sub_pix):
    co = dm.Correlation_map(img1, img2, window_size=ws)
    co()
    return co, dm.Matching(co, sub_pix=sub_pix)()


@pytest.mark.parametrize('name', ['tile_16x16_ws5_noise', 'tile_16x16_ws5_sine', 'tile_8x32_ws3', 'tile_32x8_ws5',
                                  'tile_16x16_ws15', 'tile_32x32_ws5'])
def test_end_to_end_tile_vs_reference(dm, name):
    g = load_golden(name)
    _, got = _end_to_end_tile(dm, g['img1'], g['img2'], int(g['ws']), False)
    ref = g['map_nosub']
    bad = np.mean((got[:2] != ref[:2]).any(0))
    assert bad <= max(MAX_INDEX_DISAGREEMENT, 1.0 / ref[0].size), bad
    ok = (got[:2] == ref[:2]).all(0)
    assert np.allclose(got[2][ok], ref[2][ok], rtol=LEVEL_RTOL, atol=LEVEL_ATOL)
    _, got = _end_to_end_tile(dm, g['img1'], g['img2'], int(g['ws']), True)
    ref = g['map_sub']
    close = np.abs(got[:2] - ref[:2]).max(0) <= 1e-3 * np.maximum(1.0, np.abs(ref[:2]).max(0))
    assert 1.0 - close.mean() <= max(MAX_INDEX_DISAGREEMENT, 1.0 / ref[0].size)


def test_flat_patch_gives_nan_like_reference(dm):
    g = load_golden('tile_8x8_ws5_flat')
    co, got = _end_to_end_tile(dm, g['img1'], g['img2'], int(g['ws']), True)
    assert np.isnan(g['co_map']).any()
    assert np.array_equal(np.isnan(co.co_map), np.isnan(g['co_map']))
    assert np.array_equal(np.isnan(got), np.isnan(g['map_sub']))


@pytest.mark.parametrize('name', SOLVER_CASES)
def test_image_cut_solver_vs_reference(dm, name):
    g = load_golden(name)
    s = dm.ImageCutSolver(g['img1'], g['img2'], image_size=list(g['image_size']), stride=list(g['stride']),
                          window_size=int(g['ws']), degree_map_mode=[str(m) for m in g['modes']], sub_pix=bool(g['sub_pix']))
    d, sc = s()
    assert list(s.len) == list(g['len'])
    assert d.shape == g['d_map'].shape and sc.shape == g['out_map'].shape and d.dtype == np.float64
    assert np.mean(np.abs(d - g['d_map']) > 0.5) <= MAX_INDEX_DISAGREEMENT
    close = np.abs(d - g['d_map']) <= 1e-3 * np.maximum(1.0, np.abs(g['d_map']))
    assert close.mean() >= 1.0 - MAX_INDEX_DISAGREEMENT
    assert np.mean(np.abs(sc - g['out_map']) > 1e-3) <= MAX_INDEX_DISAGREEMENT
    # the batched path and the tile-by-tile class path are the same kernels
    s2 = dm.ImageCutSolver(g['img1'], g['img2'], image_size=list(g['image_size']), stride=list(g['stride']),
                           window_size=int(g['ws']), degree_map_mode=[str(m) for m in g['modes']], sub_pix=bool(g['sub_pix']))
    s2.log_flg = False
    s2._cut_and_pool()
    s2._execute_matching_per_tile(list(d.shape[1:]))
    _same_up_to_accumulation_rounding(s2.d_map, s2.out_map, d, sc, exact=not s.info.used_fused)


@pytest.mark.parametrize('shape,size,stride,ws,sub', [((96, 96), 16, 12, 5, True), ((200, 168), 32, 30, 5, True),
                                                      ((300, 300), 64, 60, 15, True), ((150, 230), 32, 32, 3, False)])
def test_fused_path_equals_materialising_path(dm, shape, size, stride, ws, sub):
    """The fused tcgen05 path (level 0 never in HBM) recomputes level-0 values with the same
    formula from exact dot products: planes must agree with the materialising path."""
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair(shape, seed=5, mode='sine', amp=4)
    res = []
    for fused in (0, 1):
        s = dm.ImageCutSolver(i1, i2, image_size=[size, size], stride=[stride, stride], window_size=ws,
                              degree_map_mode=['elevation', 'elevation2', 'distance'], sub_pix=sub)
        s.log_flg = False
        s.fused = fused
        d, sc = s()
        assert s.info.used_fused == fused
        res.append((d, sc))
    (d0, s0), (d1, s1) = res
    assert np.mean(np.abs(d0 - d1) > 1e-4) <= 2e-4
    assert np.mean(np.abs(s0 - s1) > 1e-5) <= 2e-4


@pytest.mark.parametrize('size,stride,ws', [((32, 64), (30, 60), 5), ((64, 32), (64, 32), 7), ((16, 64), (12, 50), 3), ((8, 32), (8, 32), 5)])
def test_fused_path_non_square_tiles(dm, size, stride, ws):
    """Non-square patch grids (short side a power of two dividing the long side) on the fused
    path vs the materialising path, and vs the oracle on one tile."""
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    shape = (size[0] * 2 + stride[0] + ws + 9, size[1] * 2 + stride[1] + ws + 5)
    i1, i2 = stereo_pair(shape, seed=17, mode='sine', amp=3)
    res = []
    for fused in (0, 1):
        s = dm.ImageCutSolver(i1, i2, image_size=list(size), stride=list(stride), window_size=ws,
                              degree_map_mode=['elevation', 'elevation2'], sub_pix=True)
        s.log_flg = False
        s.fused = fused
        d, sc = s()
        assert s.info.used_fused == fused
        res.append((d, sc))
    assert np.mean(np.abs(res[0][0] - res[1][0]) > 1e-4) <= 5e-4
    assert np.mean(np.abs(res[0][1] - res[1][1]) > 1e-5) <= 5e-4
    # every tile against the oracle: <= 0.1 % integer disagreement, each one a near-tie of the oracle's own decision
    assert_parity(scene_report(i1, i2, size, stride, ws, ['elevation', 'elevation2'], res[1][0], res[1][1], True))


def test_fused_path_flat_patch_nan(dm):
    """A flat patch poisons its ancestors with NaN in the reference; both GPU paths must
    put the NaNs in the same pixels."""
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((96, 96), seed=8, mode='shift', amp=2, plain_noise=True)
    i1 = i1.copy()
    i1[30:35, 40:45] = 77
    res = []
    for fused in (0, 1):
        s = dm.ImageCutSolver(i1, i2, image_size=[16, 16], stride=[12, 12], window_size=5, degree_map_mode=['elevation'], sub_pix=True)
        s.log_flg = False
        s.fused = fused
        res.append(s())
    (d0, s0), (d1, s1) = res
    assert np.isnan(s0).any()
    assert np.array_equal(np.isnan(d0), np.isnan(d1)) and np.array_equal(np.isnan(s0), np.isnan(s1))
    ref_d, ref_s = O.image_cut_solver(i1, i2, (16, 16), (12, 12), 5, ('elevation',), True)
    assert np.array_equal(np.isnan(ref_s), np.isnan(s1))


def test_descriptor_kernels_agree(dm):
    """The 8-patches-per-group descriptor kernel and the generic one write identical rows/stats, on
    both sides (patch image: the correction entries hold S'; search image: the three bf16 parts of
    -S'/K), and the rows match numpy."""
    import torch
    from deepmatching_stereo_matching_b200 import _native
    from deepmatching_stereo_matching_b200.synth import texture
    lib = _native.lib()
    for ws in (3, 5, 7, 9, 11, 13, 15):
        e2 = ws - 1
        sc = torch.from_numpy(texture((40 + e2, 60 + e2), seed=ws)).cuda()
        H, W = sc.shape
        origin = torch.tensor([[3, 5], [7, 20]], dtype=torch.int32, device='cuda')
        kpad = lib.dm_kpad(ws)
        rstr = 8 if ws <= 8 else 16
        slots = [ky * rstr + kx for ky in range(ws) for kx in range(ws, rstr)][:3]
        for side in (1, 2, 0):
            out = []
            for (t0, t1) in ((8, 16), (8, 12)):           # t1 % 8 == 0 -> row kernel, else generic
                desc = torch.zeros((2 * t0 * t1, kpad), dtype=torch.bfloat16, device='cuda')
                stat = torch.zeros((2 * t0 * t1 * 6,), dtype=torch.float32, device='cuda')
                _native.check(lib.dm_descriptors(_native.ptr(sc), H, W, W, _native.ptr(origin), 2, t0, t1, ws, side,
                                                 _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
                st = stat.cpu().numpy()
                inv_table = st[2 * t0 * t1 * 4:2 * t0 * t1 * 5].reshape(2, t0, t1)
                out.append((desc.float().cpu().numpy().reshape(2, t0, t1, kpad), st[:2 * t0 * t1 * 4].reshape(2, t0, t1, 4), inv_table))
            (da, sa, ia), (db, sb, ib) = out
            assert np.array_equal(da[:, :, :12], db) and np.array_equal(sa[:, :, :12], sb) and np.array_equal(ia[:, :, :12], ib)
            assert np.array_equal(ib, sb[..., 1])             # the compact column table is inv
            # and against numpy: centred window values in the k = ky*rstride + kx layout
            img = sc.cpu().numpy().astype(np.int64)
            for (n, i, j) in ((0, 0, 0), (1, 7, 11), (0, 3, 9)):
                oy, ox = origin[n].tolist()
                win = img[oy + i:oy + i + ws, ox + j:ox + j + ws]
                mean = (2 * win.sum() + ws * ws) // (2 * ws * ws)
                exp = np.zeros(kpad)
                for ky in range(ws):
                    exp[ky * rstr:ky * rstr + ws] = win[ky] - mean
                rs = int((win - mean).sum())
                got = db[n, i, j].copy()
                if side == 1:
                    assert all(got[k] == rs for k in slots)
                elif side == 2:
                    c = -np.float32(rs) / np.float32(ws * ws)
                    assert np.float32(got[slots[0]]) + np.float32(got[slots[1]]) + np.float32(got[slots[2]]) == c
                    assert abs(got[slots[1]]) <= abs(c) * 2.0 ** -8 + 1e-30 and abs(got[slots[2]]) <= abs(c) * 2.0 ** -16 + 1e-30
                got[slots] = 0
                assert np.array_equal(got, exp)
                assert sb[n, i, j, 3] == mean and sb[n, i, j, 0] == rs


def test_oracle_tile_t64_ws15(dm):
    """One tile at the C2/C3 geometry (T=64, ws=15) against the numpy oracle."""
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((78, 78), seed=21, mode='sine', amp=5)
    cm = O.correlation_map(i1, i2, 15)
    ref = O.matching(cm['co_map_list'], True)
    co, got = _end_to_end_tile(dm, i1, i2, 15, True)
    assert np.abs(co.co_map - cm['co_map']).max() <= 2e-6
    for a, b in zip(co.co_map_list, cm['co_map_list']):
        assert np.allclose(a, b, rtol=LEVEL_RTOL, atol=LEVEL_ATOL)
    bad = np.mean((np.floor(got[:2] + 0.5) != np.floor(ref[:2] + 0.5)).any(0))
    assert bad <= MAX_INDEX_DISAGREEMENT


def test_batch_of_pairs_equals_one_by_one(dm):
    """Config-4 style batch (equally sized pairs, sub_pix + post-hoc sub_pix_cal): one batched
    call must reproduce the per-pair ImageCutSolver results bit for bit, on both paths."""
    from deepmatching_stereo_matching_b200 import image_cut_solver as ics
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    pairs = [stereo_pair((150, 182), seed=100 + b, mode='sine', amp=4) for b in range(3)]
    i1 = np.stack([p[0] for p in pairs]); i2 = np.stack([p[1] for p in pairs])
    kw = dict(image_size=[32, 32], stride=[32, 32], window_size=5, degree_map_mode=['elevation', 'elevation2'], sub_pix=True)
    for fused in (1, 0):
        d, sc = ics.solve_batch(i1, i2, fused=fused, **kw)
        assert d.shape == (3, 2, 96, 128) and sc.shape == (3, 96, 128)
        for b in range(3):
            s = dm.ImageCutSolver(pairs[b][0], pairs[b][1], **kw)
            s.log_flg = False
            s.fused = fused
            db, sb = s()
            assert np.array_equal(d[b], db, equal_nan=True) and np.array_equal(sc[b], sb, equal_nan=True)
    # the batched, pipelined refinement equals the plane-by-plane one
    from deepmatching_stereo_matching_b200.sub_pix_cal import sub_pix_cal_batch
    ref_batch = np.stack([np.stack([dm.sub_pix_cal(d[b, m], sc[b], direction=dr) for m, dr in ((0, 1), (1, 0))]) for b in range(3)])
    assert np.array_equal(sub_pix_cal_batch(d, sc, [1, 0]), ref_batch, equal_nan=True)
    assert np.array_equal(sub_pix_cal_batch(np.concatenate([d] * 7), np.concatenate([sc] * 7), [1, 0]), np.concatenate([ref_batch] * 7), equal_nan=True)   # several pieces per stream
    # the post-hoc refinement of config 4 (direction rule of image_cut_solver.py:137) vs the oracle
    rd, rs = O.image_cut_solver(pairs[0][0], pairs[0][1], (32, 32), (32, 32), 5, ('elevation', 'elevation2'), True)
    got = dm.sub_pix_cal(d[0, 0], sc[0], direction=1)
    ref = O.sub_pix_cal(rd[0], rs, direction=1)
    # sub_pix_cal has no peak guard (misc/sub_pix_cal.py:39-44): dis = d - (r1 - r_) / (2 (r1 + r_ - 2 r0)) amplifies
    # a score difference eps by ~eps / curvature^2, and |shift| > 1 is thrown away.  Where the
    # curvature of the oracle's own score mosaic is not degenerate the result must agree to 1e-3;
    # the rest is counted.
    r0, r1, rm = rs[1:-1, 1:-1], rs[1:-1, 2:], rs[1:-1, :-2]
    curv = np.abs(r1 + rm - 2 * r0)
    same_int = np.abs(d[0, 0] - rd[0])[1:-1, 1:-1] < 1e-3
    well = (curv > 3e-3) & same_int & (np.abs(sc[0] - rs)[1:-1, 1:-1] < 1e-5)
    diff = np.abs(got - ref)[1:-1, 1:-1]
    ok = ~(np.isnan(diff))
    print('sub_pix_cal: %.4f of the pixels well conditioned, worst error there %.2e; overall fraction above 1e-2: %.5f' % (
        well.mean(), diff[well & ok].max(), np.mean(diff[ok] > 1e-2)))
    assert well.mean() > 0.5
    assert np.mean(diff[well & ok] > 1e-3) <= 1e-3
    assert np.mean(diff[ok] > 1e-2) <= 5e-3
    # and the bit-exact statement: the same float64 arithmetic on the GPU's own planes
    assert np.array_equal(got, O.sub_pix_cal(d[0, 0], sc[0], direction=1), equal_nan=True)


def test_chunked_scene_equals_single_chunk(dm):
    """A workspace limit that forces several chunks of tiles must not change a single bit."""
    from deepmatching_stereo_matching_b200 import image_cut_solver as ics
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((300, 420), seed=12, mode='sine', amp=4)
    kw = dict(image_size=[32, 32], stride=[30, 30], window_size=5, degree_map_mode=['elevation', 'distance'], sub_pix=True)
    out = {}
    try:
        for fused in (1, 0):
            for limit in (48 << 30, 40 << 20):
                ics.set_workspace_limit(limit)
                s = dm.ImageCutSolver(i1, i2, **kw)
                s.log_flg = False
                s.fused = fused
                d, sc = s()
                out[(fused, limit)] = (d.copy(), sc.copy(), s.info.chunk_tiles, s.info.n_tiles)
            a, b = out[(fused, 48 << 30)], out[(fused, 40 << 20)]
            assert a[2] == a[3] and b[2] < b[3]                     # one chunk vs several
            assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1], equal_nan=True)
    finally:
        ics.set_workspace_limit(48 << 30)


def test_t128_tiles_c5_geometry(dm):
    """C5 geometry (image_size 128, ws 15, stride 124): constant-shift recovery on both GPU
    paths, their agreement, and a spot check of level-0 rows against the oracle's exact ZNCC."""
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((400, 400), seed=3, mode='shift', amp=5)
    res = []
    for fused in (0, 1):
        s = dm.ImageCutSolver(i1, i2, image_size=[128, 128], stride=[124, 124], window_size=15,
                              degree_map_mode=['elevation', 'elevation2'], sub_pix=True)
        s.log_flg = False
        s.fused = fused
        d, sc = s()
        assert d.shape == (2, 252, 252) and list(s.len) == [2, 2] and s.info.levels == 8 and s.info.used_fused == fused
        assert np.mean(np.abs(d[0] - 5.0) < 0.5) > 0.93 and np.mean(np.abs(d[1]) < 0.5) > 0.93
        res.append((d, sc))
    assert np.mean(np.abs(res[0][0] - res[1][0]) > 1e-4) <= 2e-4
    assert np.mean(np.abs(res[0][1] - res[1][1]) > 1e-5) <= 2e-4
    co = dm.Correlation_map(i1[:142, :142], i2[:142, :142], window_size=15)
    co._create_atomic_patch()
    co._create_simple_initial_co_map()
    got = co._dev['co_map'][5, 7].cpu().numpy()
    patch = i1[5:20, 7:22]
    assert np.abs(got - O.feature_value(patch, i2[:142, :142])).max() <= 2e-6


def test_scene_c2_properties(dm):
    """Full C2 size (1024^2, T=64, ws=15, stride 60): constant-shift recovery,
    determinism, and strip partition == whole-scene solve (bit-exact)."""
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((1024, 1024), seed=1, mode='shift', amp=3)
    kw = dict(image_size=[64, 64], stride=[60, 60], window_size=15, degree_map_mode=['elevation', 'elevation2'], sub_pix=False)
    s = dm.ImageCutSolver(i1, i2, **kw)
    d, sc = s()
    assert d.shape == (2, 904, 904) and list(s.len) == [15, 15]
    assert np.mean(d[0] == 3.0) > 0.93 and np.mean(d[1] == 0.0) > 0.93   # the 3 leftmost columns of every tile have no match inside the tile
    d2, sc2 = dm.ImageCutSolver(i1, i2, **kw)()
    assert np.array_equal(d, d2) and np.array_equal(sc, sc2)
    merged = np.zeros_like(d)
    for lo, hi in [(0, 7), (7, 15)]:
        p = dm.ImageCutSolver(i1, i2, **kw)
        p.tile_rows = (lo, hi)
        dp, _ = p()
        merged[:, p.info.row_lo:p.info.row_hi] = dp[:, p.info.row_lo:p.info.row_hi]
    assert np.array_equal(merged, d)


def test_bilateral_filter_bit_exact(dm):
    """dm_bilateral_u8 against cv2.bilateralFilter (IPP off) on the golden plane -- the live branch of
    optimize_looper.py:76-77 -- and against the oracle on a larger random plane."""
    g = load_golden('bilateral')
    for k in g:
        if not k.startswith('out_plain_'):
            continue
        _, _, d, sc, ss = k.split('_')
        got = dm.bilateral_filter(g['img'], int(d[1:]), float(sc), float(ss))
        assert got.dtype == np.uint8 and np.array_equal(got, g[k]), k
    rng = np.random.default_rng(9)
    big = rng.integers(0, 256, size=(301, 517), dtype=np.uint8)
    big[100:200, 100:300] = 128
    for d, sc, ss in ((7, 5, 5), (5, 30.0, 2.0), (15, 80.0, 6.0)):
        assert np.array_equal(dm.bilateral_filter(big, d, sc, ss), O.bilateral_filter_u8(big, d, sc, ss))
    with pytest.raises(ValueError):
        dm.bilateral_filter(big.astype(np.float64), 7, 5, 5)


def test_window_size_sweep_on_one_context(dm):
    """Solves with different window sizes on ONE context: the cached CUDA graph of the upper-pyramid
    launches bakes in workspace offsets that depend on kpad (64 for ws 3-7, 192 / 256 above); same scene
    and tile geometry with another window must not replay it.  Each result equals a fresh context's."""
    from deepmatching_stereo_matching_b200 import image_cut_solver as ics
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((300, 300), seed=23, mode='sine', amp=5)

    def solve(ws):
        s = dm.ImageCutSolver(i1, i2, image_size=[32, 32], stride=[30, 30], window_size=ws, degree_map_mode=['elevation', 'elevation2'], sub_pix=True)
        s.log_flg = False
        s.fused = 1
        s.devices = [0]
        d, sc = s()
        return d.copy(), sc.copy()

    fresh = {}
    for ws in (15, 9, 5):
        ics._CTX.clear()
        fresh[ws] = solve(ws)
    ics._CTX.clear()
    for ws in (15, 5, 9, 15, 5):            # one context: large workspace first, then smaller needs inside it
        d, sc = solve(ws)
        assert np.array_equal(d, fresh[ws][0], equal_nan=True) and np.array_equal(sc, fresh[ws][1], equal_nan=True), ws


def test_c2_bench_workload_vs_oracle(dm):
    """The benchmarked workload itself (bench.py config c2: 1024^2 sine warp, ws 15, image_size 64,
    stride 60, sub_pix, fused path): 16 evenly spaced tiles of the mosaic against the oracle with
    bench.py's own parity block, and four of them with the near-tie classification."""
    import bench
    i1, i2 = bench.make_scene('c2')
    c = bench.CONFIGS['c2']
    s = dm.ImageCutSolver(i1, i2, image_size=[c['T']] * 2, stride=[c['stride']] * 2, window_size=c['ws'], degree_map_mode=bench.MODES, sub_pix=True)
    s.log_flg = False
    s.devices = [0]
    d, sc = s()
    assert s.info.used_fused == 1 and d.shape == (2, 904, 904)
    planes = np.concatenate([d, sc[None]], 0)[None]
    par = bench.parity_block('c2', i1, i2, planes, 16)
    print('c2 parity block: %s' % par)
    assert par['ok'], par
    assert par['int_disagreement'] <= MAX_INDEX_DISAGREEMENT and par['score_max_abs'] <= 1e-3
    assert_parity(scene_report(i1, i2, (64, 64), (60, 60), 15, bench.MODES, d, sc, True, tiles=[0, 37, 112, 224]))


@pytest.mark.timeout(1500)
def test_full_t128_tile_vs_oracle(dm):
    """One FULL tile at the C5 geometry (image_size 128, ws 15): both GPU paths against the oracle
    (float64 pyramid of 16384 x 16384 entries; the oracle evaluates its correlation in row blocks)."""
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((266, 266), seed=33, mode='sine', amp=32)      # one 128 x 128 tile (len = floor((266 - 142) / 124) = 1)
    res = []
    for fused in (1, 0):
        s = dm.ImageCutSolver(i1, i2, image_size=[128, 128], stride=[124, 124], window_size=15, degree_map_mode=['elevation', 'elevation2'], sub_pix=True)
        s.log_flg = False
        s.fused = fused
        s.devices = [0]
        d, sc = s()
        assert list(s.len) == [1, 1] and d.shape == (2, 128, 128) and s.info.used_fused == fused and s.info.levels == 8
        res.append((d.copy(), sc.copy()))
    (len0, len1), trimmed = O.tile_grid(i1.shape, (128, 128), (124, 124), 15)
    a, b = i1[:trimmed[0], :trimmed[1]], i2[:trimmed[0], :trimmed[1]]
    cm = O.correlation_map(a, b, 15)
    pre, margin = O.matching_margins(cm['co_map_list'])
    out = O.sub_pix(cm['co_map_list'][0], pre)
    del cm
    for (d, sc), label in zip(res, ('fused', 'materialising')):
        n_bad = 0
        bad_any = np.zeros((128, 128), bool)
        for m, mode in enumerate(('elevation', 'elevation2')):
            ref = O.cal_map(out, mode)
            bad = np.abs(d[m] - ref) > 0.5
            bad_any |= bad
            n_bad += int(bad.sum())
            rel = np.abs(d[m] - ref) / np.maximum(1.0, np.abs(ref))
            assert np.mean(rel[~bad] > 1e-3) <= 1e-3, label
        print('T=128 %s: %d of %d integer disparities differ, worst margin %.2e; score max abs %.2e' % (
            label, n_bad, 2 * 128 * 128, margin[bad_any].max() if bad_any.any() else 0.0, np.abs(sc - out[2])[~bad_any].max()))
        assert n_bad <= MAX_INDEX_DISAGREEMENT * 2 * 128 * 128, label
        assert not (bad_any & ~(margin < NEAR_TIE)).any(), label
        assert np.mean(np.abs(sc - out[2])[~bad_any] > 1e-3) <= 1e-3, label


def test_c4_pair_vs_oracle(dm):
    """One pair of the C4 workload (512^2, image_size 32, ws 5, stride 32, sub_pix, then sub_pix_cal on
    both planes with the direction rule of image_cut_solver.py:137): all 196 tiles against the oracle."""
    import bench
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((512, 512), seed=100, mode='sine', amp=8)
    s = dm.ImageCutSolver(i1, i2, image_size=[32, 32], stride=[32, 32], window_size=5, degree_map_mode=bench.MODES, sub_pix=True)
    s.log_flg = False
    s.devices = [0]
    d, sc = s()
    assert s.info.used_fused == 1 and d.shape == (2, 448, 448) and s.info.n_tiles == 196
    assert_parity(scene_report(i1, i2, (32, 32), (32, 32), 5, bench.MODES, d, sc, True))
    for m, direction in ((0, 1), (1, 0)):
        got = dm.sub_pix_cal(d[m], sc, direction=direction)
        assert np.array_equal(got, O.sub_pix_cal(d[m], sc, direction=direction), equal_nan=True)


@pytest.mark.parametrize('fused', [1, 0])
def test_tile_ranges_assemble_the_whole_scene(dm, fused):
    """Ranges of tiles with boundaries in the middle of tile rows (what dm_multi_* and the ranks of
    bench.py use: shares differ by at most one tile), each solved on its own into the same host
    arrays / streamed into the same device mosaic: the pieces must assemble the bits of the whole solve."""
    import torch
    from deepmatching_stereo_matching_b200 import _native
    from deepmatching_stereo_matching_b200.image_cut_solver import pinned_empty
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((420, 520), seed=19, mode='sine', amp=5)
    args = (i1.shape, [32, 32], [30, 30], 5, 'cv2.TM_CCOEFF_NORMED', ['elevation', 'distance'], True)
    whole = _native.scene_params(*args, fused=fused)
    info = _native.scene_geometry(whole)
    n_tiles = info.len0 * info.len1
    assert info.len1 > 3 and n_tiles > 40
    ctx = _native.Context()
    ref_d = pinned_empty((2, info.out_h, info.out_w), np.float64); ref_s = pinned_empty((info.out_h, info.out_w), np.float64)
    ctx.solve_host(whole, i1, i2, ref_d, ref_s)
    cuts = [0, 1, info.len1 + 3, 2 * info.len1, n_tiles // 2 + 1, n_tiles - 2, n_tiles]       # one tile, mid-row cuts, a whole-row cut, a 2-tile tail
    got_d = pinned_empty(ref_d.shape, np.float64); got_s = pinned_empty(ref_s.shape, np.float64)
    got_d[...] = np.nan; got_s[...] = np.nan
    d1, d2 = torch.from_numpy(i1).cuda(), torch.from_numpy(i2).cuda()
    local = torch.zeros((3, info.out_h, info.out_w), dtype=torch.float64, device='cuda')
    mosaic = torch.full((3, info.out_h, info.out_w), float('nan'), dtype=torch.float64, device='cuda')
    tiles = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        prm = _native.scene_params(*args, fused=fused, tiles=(a, b))
        part = _native.scene_geometry(prm)
        assert part.n_tiles == b - a and part.row_lo == 30 * (a // info.len1)
        inf = ctx.solve_host(prm, i1, i2, got_d, got_s)
        assert inf.used_fused == fused
        tiles += inf.n_tiles
        ctx.solve_stream(prm, d1, d2, local[:-1], local[-1], mosaic[:-1], mosaic[-1])
    torch.cuda.synchronize()
    assert tiles == n_tiles
    assert np.array_equal(got_d, ref_d) and np.array_equal(got_s, ref_s)
    assert np.array_equal(mosaic[:-1].cpu().numpy(), ref_d) and np.array_equal(mosaic[-1].cpu().numpy(), ref_s)
    ctx.close()


def test_row_argmax_kernel(dm):
    """dm_row_argmax (bad_matching.py:68-70) against np.argmax of every patch's own map row: first
    maximum on ties, the first NaN wins, rows shorter and longer than a warp."""
    import torch
    from deepmatching_stereo_matching_b200 import _native
    rng = np.random.default_rng(3)
    for (n, t0, t1) in [(2, 8, 16), (1, 4, 80), (3, 16, 32), (1, 2, 5)]:
        x = rng.integers(0, 6, size=(n, t0, t1, t0, t1)).astype(np.float32) / 5.0        # many exact ties
        x[0, 1, 2, 1, 3] = np.nan
        x[0, 1, 2, 1, 4] = np.nan
        x[-1, 0, 0, 0, :] = 0.25
        dev = torch.from_numpy(x).cuda()
        arg = torch.empty((n, t0, t1), dtype=torch.int32, device='cuda')
        rows = torch.empty((n, t0, t1, t1), dtype=torch.float32, device='cuda')
        _native.check(_native.lib().dm_row_argmax(_native.ptr(dev), n, t0, t1, _native.ptr(arg), _native.ptr(rows), _native.stream_ptr()))
        want_rows = np.stack([[[x[k, i, j, i, :] for j in range(t1)] for i in range(t0)] for k in range(n)])
        assert np.array_equal(rows.cpu().numpy(), want_rows, equal_nan=True)
        assert np.array_equal(arg.cpu().numpy(), np.argmax(want_rows, axis=-1))
        _native.check(_native.lib().dm_row_argmax(_native.ptr(dev), n, t0, t1, _native.ptr(arg), None, _native.stream_ptr()))
        assert np.array_equal(arg.cpu().numpy(), np.argmax(want_rows, axis=-1))


def test_final_level_kernels_agree(dm, tmp_path):
    """The final level of the fused path exists twice: one thread per patch (product) and one warp per quad
    of patches (DM_FINAL_QUAD=1, the round-1 kernel, kept as a measurement aid).  Same arithmetic and the
    same order of comparisons: the planes must agree bit for bit -- window sizes 3..15, non-square grids,
    TM_CCOEFF, a flat patch (NaN), with and without the parabola fit.  (The switch is read once per
    process, so each variant runs in its own interpreter.)"""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import deepmatching_stereo_matching_b200 as dm
from deepmatching_stereo_matching_b200.synth import stereo_pair
out = {}
for k, (shape, size, stride, ws, feat, sub) in enumerate([((200, 264), (32, 32), (30, 30), 15, 'cv2.TM_CCOEFF_NORMED', True),
                                                         ((150, 230), (16, 64), (12, 50), 3, 'cv2.TM_CCOEFF_NORMED', True),
                                                         ((120, 120), (16, 16), (14, 14), 5, 'cv2.TM_CCOEFF', True),
                                                         ((160, 160), (32, 32), (32, 32), 7, 'cv2.TM_CCOEFF_NORMED', False),
                                                         ((140, 200), (32, 32), (28, 28), 9, 'cv2.TM_CCOEFF_NORMED', True),
                                                         ((140, 140), (16, 16), (16, 16), 11, 'cv2.TM_CCOEFF_NORMED', True),
                                                         ((150, 150), (32, 32), (30, 30), 13, 'cv2.TM_CCOEFF_NORMED', True)]):
    i1, i2 = stereo_pair(shape, seed=40 + k, mode='sine', amp=4)
    if k == 2:
        i1 = i1.copy(); i1[30:35, 40:45] = 77           # a flat patch
    s = dm.ImageCutSolver(i1, i2, image_size=list(size), stride=list(stride), window_size=ws, feature_name=feat,
                          degree_map_mode=['elevation', 'elevation2', 'distance'], sub_pix=sub)
    s.log_flg = False; s.fused = 1; s.devices = [0]
    d, sc = s()
    assert s.info.used_fused == 1
    out['d%%d' %% k] = d; out['s%%d' %% k] = sc
np.savez(sys.argv[1], **out)
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    res = []
    for name, env in (('patch', {}), ('quad', {'DM_FINAL_QUAD': '1'})):
        path = str(tmp_path / (name + '.npz'))
        e = dict(os.environ); e.pop('DM_FINAL_QUAD', None); e.update(env)
        r = subprocess.run([sys.executable, '-c', code, path], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(np.load(path))
    a, b = res
    assert sorted(a.files) == sorted(b.files) and len(a.files) == 14
    assert np.isnan(a['s2']).any()
    for k in a.files:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_launch_and_store_variants_agree(dm, tmp_path):
    """Round 2's second session changed HOW several stages launch, stage and store -- never what they compute: the pooled
    epilogue's pair flush (image_size 64) and 32-float staging rows (128), the two-copy upper aggregation, the
    bulk tensor stores those regions leave by, the persistent first aggregation of small maps, both images' descriptors in one launch (and the 8-lane window sums by
    shuffles), programmatic dependent launch.  Each has a switch back to the form it replaced;
    the planes of four scenes -- tiles of 16, 32, 64 and 128 -- must agree bit for bit between the product, all switches
    thrown, and the attribute on every launch of the chain.  (The switches are read once per process.)"""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import deepmatching_stereo_matching_b200 as dm
from deepmatching_stereo_matching_b200.synth import stereo_pair
out = {}
for k, (shape, size, stride, ws) in enumerate([((150, 150), (16, 16), (14, 14), 5), ((260, 300), (32, 32), (32, 32), 5),
                                                ((340, 400), (64, 64), (60, 60), 15), ((400, 400), (128, 128), (124, 124), 15)]):
    i1, i2 = stereo_pair(shape, seed=70 + k, mode='sine', amp=size[0] // 8)
    s = dm.ImageCutSolver(i1, i2, image_size=list(size), stride=list(stride), window_size=ws,
                          degree_map_mode=['elevation', 'elevation2'], sub_pix=True)
    s.log_flg = False; s.fused = 1; s.devices = [0]
    d, sc = s()
    assert s.info.used_fused == 1
    out['d%%d' %% k] = d; out['s%%d' %% k] = sc
np.savez(sys.argv[1], **out)
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    switches = ['DM_CORR_NO_PAIR_FLUSH', 'DM_CORR_NO_WIDE', 'DM_CORR_NO_TMA_STORE', 'DM_AGG_NO_MERGE', 'DM_FIRST_CTA', 'DM_DESC_SPLIT', 'DM_PDL']
    res = []
    for name, env in (('product', {}), ('replaced', {'DM_CORR_NO_PAIR_FLUSH': '1', 'DM_CORR_NO_WIDE': '1', 'DM_AGG_NO_MERGE': '1',
                                                     'DM_FIRST_CTA': '1', 'DM_DESC_SPLIT': '1', 'DM_PDL': '0'}),
                      ('flush_by_loads_and_stores', {'DM_CORR_NO_TMA_STORE': '1'}), ('pdl_everywhere', {'DM_PDL': '31'})):
        path = str(tmp_path / (name + '.npz'))
        e = dict(os.environ)
        for k in switches:
            e.pop(k, None)
        e.update(env)
        r = subprocess.run([sys.executable, '-c', code, path], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(np.load(path))
    a = res[0]
    assert len(a.files) == 8
    for other in res[1:]:
        assert sorted(a.files) == sorted(other.files)
        for k in a.files:
            assert np.array_equal(a[k], other[k], equal_nan=True), k


def test_gauss_seidel_loops_bit_exact(dm):
    """The reference's sequential in-place smoothing loops (its `if 0:` branch, optimize_looper.py:55-74) through the
    `misc.*` import paths: misc/optimize_loop.py::optimize_loop and misc/opt_loop.py::optimize_loop_bilateral_*
    against the live reference's outputs -- same bits, including the sequentially accumulated error -- and
    make_weight within an ulp (CUDA's exp against numpy's)."""
    from misc.optimize_loop import optimize_loop
    from misc.opt_loop import make_weight, optimize_loop_bilateral_horizon, optimize_loop_bilateral_vertical
    g = load_golden('gauss_seidel')
    size = g['d'].shape
    for k in range(3):
        alpha, exclusion, loops = g['ol%d_cfg' % k]
        x = g['d'].copy()
        errs = []
        for _ in range(int(loops)):
            y, err = optimize_loop(x, g['co'], float(alpha), int(exclusion), size)
            assert y is not x                        # the reference clamps into a new array
            x = y
            errs.append(err)
        assert np.array_equal(x, g['ol%d_out' % k]), k
        assert np.array_equal(np.array(errs), g['ol%d_err' % k]), k
    for k in range(2):
        s0, s1, exclusion, loops = g['bl%d_cfg' % k]
        sigma = np.array([int(s0), int(s1)])
        gw, cw = make_weight(g['d'], int(exclusion), size, sigma)
        assert gw.shape == g['bl%d_gw' % k].shape and cw.shape == g['bl%d_cw' % k].shape
        assert np.allclose(gw, g['bl%d_gw' % k], rtol=4e-16, atol=0) and np.allclose(cw, g['bl%d_cw' % k], rtol=4e-16, atol=0)
        assert np.array_equal(cw == 0, g['bl%d_cw' % k] == 0)
        for name, fn in (('h', optimize_loop_bilateral_horizon), ('v', optimize_loop_bilateral_vertical)):
            x = g['d'].copy()
            errs = []
            for _ in range(int(loops)):
                y, err = fn(x, g['bl%d_cw' % k], g['bl%d_gw' % k], g['co'], 0.008, int(exclusion), size)
                assert y is x                        # in place, like the reference
                errs.append(err)
            assert np.array_equal(x, g['bl%d_%s_out' % (k, name)]), (k, name)
            assert np.array_equal(np.array(errs), g['bl%d_%s_err' % (k, name)]), (k, name)
            # with the library's own weights: the same to rounding
            x2, _ = fn(g['d'].copy(), cw, gw, g['co'], 0.008, int(exclusion), size)
            x3, _ = fn(g['d'].copy(), g['bl%d_cw' % k], g['bl%d_gw' % k], g['co'], 0.008, int(exclusion), size)
            assert np.allclose(x2, x3, rtol=1e-12, atol=1e-12)
    # a larger plane against the oracle (a window sum of 13 x 13 = 169 > 128 terms takes numpy's recursive split)
    rng = np.random.default_rng(3)
    d = rng.normal(size=(70, 90)) * 2 + 4
    co = rng.random((70, 90)) + 0.3
    a, ea = optimize_loop(d, co, 0.01, 2, d.shape)
    b, eb = O.optimize_loop(d, co, 0.01, 2, d.shape)
    assert np.array_equal(a, b) and ea == eb
    gw, cw = O.make_weight(d, 6, d.shape, np.array([3, 4]))
    a, ea = optimize_loop_bilateral_vertical(d.copy(), cw, gw, co, 0.008, 6, d.shape)
    b, eb = O.optimize_loop_bilateral(d.copy(), cw, gw, co, 0.008, 6, d.shape, vertical=True)
    assert np.array_equal(a, b) and ea == eb


@pytest.mark.parametrize('kind', ['plain_noise', 'two_level'])
def test_correction_slots_on_high_contrast_input(dm, kind):
    """The tensor-core engine adds the three correction terms (-S1' * parts of S2'/K) to an accumulator that
    can be as large as 2^23..2^24 on high-contrast input (uniform noise; a random 0 / 255 pattern where
    |a'| reaches 255): co_map must stay within 2e-6 of the exact float64 oracle there as well, and the
    tensor-core and CUDA-core engines within 1e-6 of each other (min-maxed values)."""
    from deepmatching_stereo_matching_b200.synth import texture
    rng = np.random.default_rng(77)
    if kind == 'plain_noise':
        i1 = texture((78, 78), seed=91, plain_noise=True)
        i2 = texture((78, 78), seed=92, plain_noise=True)
    else:
        i1 = (rng.integers(0, 2, size=(78, 78)) * 255).astype(np.uint8)
        i2 = i1.copy()
        flip = rng.random((78, 78)) < 0.2
        i2[flip] = 255 - i2[flip]
    exact = O.initial_co_map(i1, i2, 15)
    got = {}
    for engine in (0, 1):                                # 0 = auto (tcgen05 at P = 4096), 1 = CUDA cores
        co = dm.Correlation_map(i1, i2, window_size=15)
        co._create_atomic_patch()
        co._create_simple_initial_co_map(engine=engine)
        got[engine] = np.asarray(co.co_map)
        assert np.abs(got[engine] - exact).max() <= 2e-6, (kind, engine)
    assert np.abs(got[0] - got[1]).max() <= 1e-6
