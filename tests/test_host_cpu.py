"""CPU-only checks: the C-ABI library loads and exports what include/dmstereo.h declares,
geometry, argument validation of the reference-shaped classes, strip partition."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REPO
from oracle import dm_oracle as O


def test_library_exports_every_declared_symbol():
    from deepmatching_stereo_matching_b200 import _native
    header = open(os.path.join(REPO, 'include', 'dmstereo.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(dm_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 20
    lib = _native.lib()
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), 'libdmstereo.so does not export %s' % name
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert lib.dm_version() >= 100
    assert lib.dm_kpad(15) == 256 and lib.dm_kpad(5) == 64 and lib.dm_kpad(3) == 64


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """include/dmstereo.h is the drop-in boundary: it must compile as C (not only C++), and a C
    program that calls only host-side entry points must link against libdmstereo.so and run
    without a GPU (geometry, kpad, the error path of dm_correlation_set_pair_mode)."""
    import shutil
    import subprocess
    from deepmatching_stereo_matching_b200 import _native
    _native.lib()
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('gcc not available')
    src = tmp_path / 'abi.c'
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "dmstereo.h"
int main(void) {
    dm_scene_params p;
    dm_scene_info info;
    memset(&p, 0, sizeof p);
    p.scene_h = 1024; p.scene_w = 1024; p.t0 = 64; p.t1 = 64; p.s0 = 60; p.s1 = 60; p.ws = 15;
    p.method = DM_TM_CCOEFF_NORMED; p.n_modes = 1; p.modes[0] = DM_MODE_ELEVATION; p.sub_pix = 1; p.fused = -1;
    if (dm_scene_geometry(&p, &info) != DM_OK) { printf("geometry failed: %s\n", dm_last_error()); return 1; }
    if (dm_correlation_set_pair_mode(9) != DM_ERR_INVALID) return 2;
    if (dm_correlation_set_pair_mode(-1) != DM_OK) return 3;
    printf("%d %d %d %d %d %d\n", info.len0, info.len1, info.out_h, info.out_w, info.n_tiles, dm_kpad(15));
    return 0;
}
''')
    exe = tmp_path / 'abi'
    libdir = os.path.dirname(_native.LIB_PATH)
    cmd = [gcc, '-std=c99', '-Wall', '-Werror', '-I', os.path.join(REPO, 'include'), str(src), '-o', str(exe),
           '-L', libdir, '-ldmstereo', '-Wl,-rpath,' + libdir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [15, 15, 904, 904, 225, 256]


def test_pair_mode_knob_validates_its_argument():
    """dm_correlation_set_pair_mode is host-only state: -1 / 1 = CTA pairs when the shape allows,
    0 = one CTA per work item; anything else is DM_ERR_INVALID with a message."""
    from deepmatching_stereo_matching_b200 import _native
    lib = _native.lib()
    for mode in (0, 1, -1):
        assert lib.dm_correlation_set_pair_mode(mode) == 0
    assert lib.dm_correlation_set_pair_mode(7) < 0
    assert b'pair_mode' in lib.dm_last_error()
    assert lib.dm_correlation_set_pair_mode(-1) == 0


def test_struct_layout_matches_header():
    from deepmatching_stereo_matching_b200 import _native
    assert ctypes.sizeof(_native.SceneParams) == 4 * (9 + 4 + 4 + 3 + 2)
    assert ctypes.sizeof(_native.SceneInfo) == 4 * 12


@pytest.mark.parametrize('shape,size,stride,ws', [((256, 256), (32, 32), (32, 32), 5), ((1024, 1024), (64, 64), (60, 60), 15),
                                                   ((4096, 4096), (64, 64), (60, 60), 15), ((8192, 8192), (128, 128), (124, 124), 15),
                                                   ((80, 112), (16, 16), (16, 16), 3), ((100, 300), (8, 32), (8, 30), 5)])
def test_geometry_matches_oracle(shape, size, stride, ws):
    from deepmatching_stereo_matching_b200 import _native
    info = _native.scene_geometry(_native.scene_params(shape, size, stride, ws, 'cv2.TM_CCOEFF_NORMED', ['elevation'], True))
    ln, _ = O.tile_grid(shape, size, stride, ws)
    assert [info.len0, info.len1] == ln
    assert info.out_h == stride[0] * (ln[0] - 1) + size[0] and info.out_w == stride[1] * (ln[1] - 1) + size[1]
    assert info.n_tiles == ln[0] * ln[1] and info.n_map == min(size) and info.levels == int(np.log2(min(size))) + 1
    assert (info.row_lo, info.row_hi) == (0, info.out_h)


def test_geometry_table_of_survey():
    from deepmatching_stereo_matching_b200 import _native
    for shape, size, stride, ws, tiles, out in [((256, 256), 32, 32, 5, 36, 192), ((1024, 1024), 64, 60, 15, 225, 904),
                                                 ((4096, 4096), 64, 60, 15, 4356, 3964), ((8192, 8192), 128, 124, 15, 4096, 7940),
                                                 ((512, 512), 32, 32, 5, 196, 448)]:
        info = _native.scene_geometry(_native.scene_params(shape, (size, size), (stride, stride), ws, 'cv2.TM_CCOEFF_NORMED', ['elevation'], True))
        assert info.n_tiles == tiles and info.out_h == out


def test_geometry_rejects_bad_arguments():
    from deepmatching_stereo_matching_b200 import _native
    ok = dict(shape=(256, 256), image_size=(32, 32), stride=(32, 32), window_size=5, feature_name='cv2.TM_CCOEFF_NORMED', modes=['elevation'], sub_pix=True)
    for bad in [dict(window_size=4), dict(image_size=(12, 12)), dict(image_size=(8, 20)), dict(shape=(30, 30)), dict(stride=(0, 32))]:
        kw = dict(ok); kw.update(bad)
        with pytest.raises(_native.DmError):
            _native.scene_geometry(_native.scene_params(**kw))
    prm = _native.scene_params(**dict(ok, tile_rows=(3, 9)))
    with pytest.raises(_native.DmError):
        _native.scene_geometry(prm)       # only 6 tile rows


def test_strip_geometry():
    from deepmatching_stereo_matching_b200 import _native
    from deepmatching_stereo_matching_b200.strips import partition_tile_rows, strip_rows, input_rows
    assert [hi - lo for lo, hi in partition_tile_rows(66, 8)] == [9, 9, 8, 8, 8, 8, 8, 8]
    assert partition_tile_rows(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    parts = partition_tile_rows(66, 8)
    rows = [strip_rows(lo, hi, 66, 60, 64) for lo, hi in parts]
    assert rows[0][0] == 0 and rows[-1][1] == 3964
    assert all(rows[i][1] == rows[i + 1][0] for i in range(7))
    for (lo, hi), rr in zip(parts, rows):
        info = _native.scene_geometry(_native.scene_params((4096, 4096), (64, 64), (60, 60), 15, 'cv2.TM_CCOEFF_NORMED', ['elevation'], True, (lo, hi)))
        assert (info.row_lo, info.row_hi) == rr and info.n_tiles == (hi - lo) * 66
    a, b = input_rows(9, 18, 60, 64, 15)
    assert (a, b) == (540, 60 * 17 + 78)


def test_reference_shaped_validation(capsys):
    import deepmatching_stereo_matching_b200 as dm
    with pytest.raises(SystemExit):
        dm.Feature_value('cv2.TM_SQDIFF')
    with pytest.raises(SystemExit):
        dm.Correlation_map(np.zeros((10, 10), np.uint8), np.zeros((10, 11), np.uint8))
    with pytest.raises(SystemExit):
        dm.Matching(object())
    class Co:
        co_map_list = []
    with pytest.raises(AssertionError):
        dm.Matching(Co(), filtering_mode='mean')
    with pytest.raises(AssertionError):
        dm.ImageCutSolver(np.zeros((64, 64), np.uint8), np.zeros((64, 65), np.uint8))
    with pytest.raises(SystemExit):
        dm.Calc_difference.cal_map(np.zeros((3, 4, 4)), 'nope')
    s = dm.ImageCutSolver(np.zeros((256, 256), np.uint8), np.zeros((256, 256), np.uint8))
    assert s.len == [6, 6] and s.trimed_size == [36, 36] and s.exclusive_pix == 2
    s._cut_and_pool()
    assert len(s.img_index) == 36 and s.img_index[1] == [1, 0] and s.img1_sub[0].shape == (36, 36)
    m = dm.Matching(Co())
    assert (m.filter_window_size, m.filtering, m.filtering_num, m.filtering_mode, m.sub_pix) == (3, False, 3, 'median', True)
    assert np.array_equal(dm.image_threshold(np.array([-5., 0.5, 20.])), np.array([0., 0.5, 10.]))


def test_no_tile_fits_raises_index_error_like_reference():
    import deepmatching_stereo_matching_b200 as dm
    s = dm.ImageCutSolver(np.zeros((40, 200), np.uint8), np.zeros((40, 200), np.uint8), image_size=[32, 32], stride=[32, 32], window_size=5)
    assert s.len[0] == 0
    with pytest.raises(IndexError):
        s()                                   # img_index[-1] on an empty list (image_cut_solver.py:150)
    with pytest.raises(NotImplementedError):
        dm.ImageCutSolver(np.zeros((64, 64), np.uint8), np.zeros((64, 64), np.uint8), padding=True)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import deepmatching_stereo_matching_b200 as dm
    from deepmatching_stereo_matching_b200._native import DmError
    img = np.random.default_rng(0).integers(0, 255, (20, 20), dtype=np.uint8)
    with pytest.raises(DmError):
        dm.Correlation_map(img, img, 5)()
    with pytest.raises(DmError):
        dm.ImageCutSolver(np.zeros((256, 256), np.uint8), np.zeros((256, 256), np.uint8))()
    with pytest.raises(DmError):
        dm.sub_pix_cal(np.zeros((5, 5)), np.zeros((5, 5)))


def test_misc_shims_resolve_to_the_package():
    import misc.Correlation_map, misc.Matching, misc.image_cut_solver, misc.Feature_value, misc.raw_read
    import misc.sub_pix_cal, misc.Calc_difference, misc.optimize_loop
    import deepmatching_stereo_matching_b200 as dm
    assert misc.Correlation_map.Correlation_map is dm.Correlation_map
    assert misc.image_cut_solver.ImageCutSolver is dm.ImageCutSolver
    assert misc.Matching.Matching is dm.Matching and misc.sub_pix_cal.sub_pix_cal is dm.sub_pix_cal
    assert misc.raw_read.RawRead is dm.RawRead and misc.optimize_loop.image_threshold is dm.image_threshold


def test_raw_read(tmp_path):
    import deepmatching_stereo_matching_b200 as dm
    a = (np.arange(12 * 10) * 3 % 256).astype(np.uint8).reshape(10, 12)
    p = tmp_path / 'x.raw'
    a.tofile(p)
    assert np.array_equal(dm.RawRead.read(str(p), size=(12, 10)), O.raw_read(str(p), size=(12, 10)))
    assert np.array_equal(dm.RawRead.read(str(p), size=(12, 10), rate=2), O.raw_read(str(p), size=(12, 10), rate=2))


def test_partition_rule_is_the_same_in_c_and_python():
    from deepmatching_stereo_matching_b200 import _native
    from deepmatching_stereo_matching_b200.strips import partition_tile_rows
    for len0, n in [(66, 8), (64, 8), (15, 4), (3, 8), (1, 1), (33, 2)]:
        assert _native.partition_tile_rows(len0, n) == partition_tile_rows(len0, n)
    assert _native.partition_tile_rows(66, 8)[:3] == [(0, 9), (9, 18), (18, 26)]


def test_bench_arms_print_the_same_config():
    import bench
    for name in bench.CONFIGS:
        for n in (1, 8):
            a, b = bench.config_dict(name, n), bench.config_dict(name, n)
            assert a == b and a['name'] == name and a['workload'] == bench.workload_name(name)
    assert bench.config_dict('c3', 8)['tiles'] == 4356 and bench.config_dict('c5', 8)['output'] == [1, 7940, 7940]
    assert bench.config_dict('c4', 2)['tiles'] == 12544


def test_ncu_traffic_table_is_what_bench_reads():
    """bench.py takes roofline.traffic from profiles/ncu_traffic.json: per launch of the C2 step, and per tile of the
    correlation kernel for the other shapes; a broken table would take the bench line with it."""
    import json
    import bench
    t = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json')))
    for k in ('dm_correlation_umma_kernel', 'dm_aggregate_first_kernel', 'dm_aggregate_kernel', 'dm_descriptor_row_kernel'):
        assert t['fused'][k] > 0
    for name, c in bench.CONFIGS.items():
        v = t['correlation_per_tile'].get('t%d_ws%d' % (c['T'], c['ws']))
        assert v is not None and v > 0, name
        # never below the algorithmic bytes of a tile: the pooled map written once, both descriptor blocks read once
        P = c['T'] ** 2
        kpad = 64 if c['ws'] <= 7 else 256
        assert v >= 4.0 * P * P / 4 + 2 * P * kpad * 2 * 0.5


def test_oracle_margins_follow_the_matching():
    """matching_margins walks the same path as matching() and its margins are positive distances."""
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    i1, i2 = stereo_pair((24, 40), seed=5, mode='sine', amp=2)
    cm = O.correlation_map(i1, i2, 5)
    mp, mg = O.matching_margins(cm['co_map_list'])
    assert np.array_equal(mp, O.matching(cm['co_map_list'], False), equal_nan=True)
    assert mg.shape == mp.shape[1:] and (mg > 0).all()
    r = O.match_template_matrix(i1, i2, 5)
    assert np.array_equal(r, O.match_template_matrix(i1, i2, 5, row_chunk=37))


def test_tile_range_geometry():
    from deepmatching_stereo_matching_b200 import _native
    args = ((4096, 4096), (64, 64), (60, 60), 15, 'cv2.TM_CCOEFF_NORMED', ['elevation'], True)
    parts = _native.partition_tile_rows(66 * 66, 8)
    assert [b - a for a, b in parts] == [545, 545, 545, 545, 544, 544, 544, 544]
    for a, b in parts:
        info = _native.scene_geometry(_native.scene_params(*args, tiles=(a, b)))
        assert info.n_tiles == b - a and info.row_lo == 60 * (a // 66)
        assert info.row_hi == (3964 if (b - 1) // 66 == 65 else 60 * ((b - 1) // 66 + 1))
    for bad in [(5, 5), (-1, 4), (0, 66 * 66 + 1)]:
        with pytest.raises(_native.DmError):
            _native.scene_geometry(_native.scene_params(*args, tiles=bad))
    with pytest.raises(_native.DmError):        # a tile range and a strip of tile rows exclude each other
        _native.scene_geometry(_native.scene_params(*args, tile_rows=(0, 3), tiles=(0, 10)))


@pytest.mark.parametrize('shape,size,stride,ws', [((200, 264), (32, 32), (30, 30), 5), ((150, 230), (16, 64), (12, 50), 3),
                                                   ((300, 300), (32, 32), (40, 36), 7), ((140, 140), (16, 16), (16, 16), 3)])
def test_owned_rectangles_cover_exactly_what_the_paste_order_leaves(shape, size, stride, ws):
    """dm_owned_rectangles against a brute-force paste in the reference's order (j outer, i inner,
    later tiles overwrite, misc/image_cut_solver.py:165-175): for every way of cutting the tiles into
    contiguous ranges the rectangles of a range are exactly the pixels whose final owner lies in it."""
    from deepmatching_stereo_matching_b200 import _native
    args = (shape, size, stride, ws, 'cv2.TM_CCOEFF_NORMED', ['elevation'], True)
    info = _native.scene_geometry(_native.scene_params(*args))
    n = info.len0 * info.len1
    owner = np.full((info.out_h, info.out_w), -1, dtype=np.int64)
    for j in range(info.len1):
        for i in range(info.len0):
            owner[stride[0] * i:stride[0] * i + size[0], stride[1] * j:stride[1] * j + size[1]] = i * info.len1 + j
    rng = np.random.default_rng(n)
    for parts in (1, 2, 3, 5, 8):
        cuts = [0] + sorted(rng.choice(np.arange(1, n), size=min(parts, n) - 1, replace=False).tolist()) + [n]
        seen = np.zeros_like(owner)
        for a, b in zip(cuts[:-1], cuts[1:]):
            rects = _native.owned_rectangles(_native.scene_params(*args, tiles=(a, b)))
            assert 1 <= len(rects) <= 3
            mask = np.zeros(owner.shape, bool)
            for r0, r1, c0, c1 in rects:
                assert 0 <= r0 < r1 <= info.out_h and 0 <= c0 < c1 <= info.out_w
                assert not mask[r0:r1, c0:c1].any()                  # the rectangles of a range are disjoint
                mask[r0:r1, c0:c1] = True
            want = (owner >= a) & (owner < b)
            # where the stride exceeds the tile there are pixels no tile writes (np.empty in the reference, zeros
            # here): the rectangles carry them along, what counts is the owned pixels
            assert np.array_equal(mask & (owner >= 0), want), (a, b)
            seen += mask
        assert (seen == 1).all()                                     # the ranges' rectangles tile the mosaic: every pixel exactly once


def test_tile_ranges_paste_like_the_whole_scene():
    """The oracle restricted to ranges of tiles, each range pasted through its owned rectangles into one
    mosaic, equals the whole-scene solve: the host-side contract of the multi-GPU shares."""
    from conftest import load_golden
    from deepmatching_stereo_matching_b200 import _native
    g = load_golden('solver_96_t16_s12_ws5')
    size, stride, ws = tuple(int(x) for x in g['image_size']), tuple(int(x) for x in g['stride']), int(g['ws'])
    modes = tuple(str(m) for m in g['modes'])
    args = (g['img1'].shape, size, stride, ws, 'cv2.TM_CCOEFF_NORMED', list(modes), bool(g['sub_pix']))
    info = _native.scene_geometry(_native.scene_params(*args))
    parts = _native.partition_tile_rows(info.len0 * info.len1, 3)
    d_all = np.full(g['d_map'].shape, np.nan); s_all = np.full(g['out_map'].shape, np.nan)
    for a, b in parts:
        d, s = O.image_cut_solver(g['img1'], g['img2'], size, stride, ws, modes, bool(g['sub_pix']), tiles=(a, b))
        for r0, r1, c0, c1 in _native.owned_rectangles(_native.scene_params(*args, tiles=(a, b))):
            d_all[:, r0:r1, c0:c1] = d[:, r0:r1, c0:c1]
            s_all[r0:r1, c0:c1] = s[r0:r1, c0:c1]
    d_ref, s_ref = O.image_cut_solver(g['img1'], g['img2'], size, stride, ws, modes, bool(g['sub_pix']))
    assert np.array_equal(d_all, d_ref, equal_nan=True) and np.array_equal(s_all, s_ref, equal_nan=True)
