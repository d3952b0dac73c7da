# -*- coding: utf-8 -*-
"""
CPU oracle for the DeepMatching-for-stereo hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a vectorised numpy restatement of the reference algorithm.  It is the
checker the CUDA path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package (``deepmatching_stereo_matching_b200``) never does and has no CPU
fallback.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the *live, unmodified* reference generated in the
build container by ``tests/golden/make_golden.py`` (committed ``tests/golden/*.npz``)
and, where OpenCV is importable, directly against ``cv2.matchTemplate``.

Every function cites the reference lines it restates (paths relative to the reference
repository root).  The correlation arithmetic itself lives in a third-party dependency
of the reference -- OpenCV ``cv2.matchTemplate`` (pinned ``opencv==3.4.1`` in
``environment.yml:7``; 4.13.0 in this image), file
``modules/imgproc/src/templmatch.cpp::common_matchTemplate`` -- which is restated here
from its published algorithm: exact cross-correlation, integral-image window sums in
float64, the ``|num| < t`` / ``1.125 t`` clamp rule, float32 result.
"""

import numpy as np

LAM = 1.4                     # misc/Correlation_map.py:41  (self.lam)
NEAR_ZERO = 0.0001            # misc/Matching.py:74

TM_CCOEFF = 4                 # cv2.TM_CCOEFF
TM_CCOEFF_NORMED = 5          # cv2.TM_CCOEFF_NORMED
FEATURE_NAMES = {'cv2.TM_CCOEFF_NORMED': TM_CCOEFF_NORMED, 'cv2.TM_CCOEFF': TM_CCOEFF}  # misc/Feature_value.py:24

_FLT_EPSILON = float(np.finfo(np.float32).eps)
_DBL_EPSILON = float(np.finfo(np.float64).eps)


# --------------------------------------------------------------------------------------
# descriptors
# --------------------------------------------------------------------------------------
def atomic_patches(img, ws):
    """misc/Correlation_map.py:51-67 -- every ws x ws window, (T0,T1,ws,ws) uint8."""
    img = np.asarray(img)
    win = np.lib.stride_tricks.sliding_window_view(img, (ws, ws))
    return np.ascontiguousarray(win).astype(np.uint8)


def _im2col(img, ws):
    """(T0*T1, ws*ws) float64 matrix of the windows of ``img`` (row-major positions)."""
    win = np.lib.stride_tricks.sliding_window_view(np.asarray(img), (ws, ws))
    t0, t1 = win.shape[:2]
    return win.reshape(t0 * t1, ws * ws).astype(np.float64), (t0, t1)


# --------------------------------------------------------------------------------------
# correlation  (OpenCV matchTemplate restated)  +  min-max
# --------------------------------------------------------------------------------------
def match_template_matrix(img, template, ws, method=TM_CCOEFF_NORMED, row_chunk=None):
    """All patches of ``img`` against all windows of ``template`` at once.

    Restates ``cv2.matchTemplate(patch, template, method)`` (call site
    misc/Feature_value.py:41) for every atomic patch: returns a float32 matrix
    ``R[p, q]`` with p = patch index in ``img`` and q = window index in ``template``.
    OpenCV algorithm (templmatch.cpp::common_matchTemplate): ``num = sum(T*I) -
    mean(T)*sum(I)``; for the normed variant ``t = sqrt(max(sum(I^2) - sum(I)^2/K, 0)) *
    sqrt(K)*std(T)`` and ``num/t`` if ``|num| < t``, ``+-1`` if ``|num| < 1.125 t``, else
    0; a flat template yields an all-ones map; result stored as float32.
    """
    a1, _ = _im2col(img, ws)
    if a1.shape[0] > 4096 and row_chunk is None:
        row_chunk = 2048                             # a 128 x 128 grid: 2 GB per float64 temporary otherwise
    if row_chunk is not None and a1.shape[0] > row_chunk:
        # the rows (patches) are independent: evaluate them in blocks to bound the temporaries
        win = np.lib.stride_tricks.sliding_window_view(np.asarray(img), (ws, ws))
        t1 = win.shape[1]
        out = []
        for r0 in range(0, a1.shape[0], row_chunk):
            out.append(_match_rows(a1[r0:r0 + row_chunk], template, ws, method))
        return np.concatenate(out, 0)
    return _match_rows(a1, template, ws, method)


def _match_rows(a1, template, ws, method):
    """match_template_matrix for the patch rows ``a1`` (n, ws*ws) float64."""
    a2, _ = _im2col(template, ws)
    k = float(ws * ws)
    inv_area = 1.0 / k
    cc = a1 @ a2.T                                   # exact: integer sums < 2^53
    s1 = a1.sum(1)
    q1 = (a1 * a1).sum(1)
    s2 = a2.sum(1)
    q2 = (a2 * a2).sum(1)
    templ_mean = s1 * inv_area
    num = cc - templ_mean[:, None] * s2[None, :]
    if method == TM_CCOEFF:
        return num.astype(np.float32)
    # meanStdDev of the template (population std)
    templ_var = np.maximum(q1 * inv_area - templ_mean * templ_mean, 0.0)
    templ_sdv = np.sqrt(templ_var)
    templ_norm2 = templ_sdv * templ_sdv
    flat_templ = templ_norm2 < _DBL_EPSILON
    templ_norm = np.sqrt(templ_norm2) / np.sqrt(inv_area)
    wnd_mean2 = s2 * s2 * inv_area
    wnd_sum2 = q2
    diff2 = np.maximum(wnd_sum2 - wnd_mean2, 0.0)
    t_w = np.where(diff2 <= np.minimum(0.5, 10.0 * _FLT_EPSILON * wnd_sum2), 0.0, np.sqrt(diff2))
    t = templ_norm[:, None] * t_w[None, :]
    absn = np.abs(num)
    with np.errstate(divide='ignore', invalid='ignore'):
        res = np.where(absn < t, num / t, np.where(absn < t * 1.125, np.sign(num), 0.0))
    res[flat_templ, :] = 1.0
    return res.astype(np.float32)


def min_max(x):
    """misc/Feature_value.py:32-37 -- (x - min) / (max - min) in the array's dtype."""
    with np.errstate(divide='ignore', invalid='ignore'):
        mn = x.min(axis=-1, keepdims=True)
        mx = x.max(axis=-1, keepdims=True)
        return (x - mn) / (mx - mn)


def feature_value(patch, image, method=TM_CCOEFF_NORMED):
    """misc/Feature_value.py:39-43 for one patch: float32 map, min-maxed over the map.

    ``patch`` and ``image`` are any sizes with patch <= image (for_igarss/cor_map.py:33-35).
    """
    patch = np.asarray(patch)
    image = np.asarray(image)
    ph, pw = patch.shape
    win = np.lib.stride_tricks.sliding_window_view(image, (ph, pw))
    oh, ow = win.shape[:2]
    a2 = win.reshape(oh * ow, ph * pw).astype(np.float64)
    a1 = patch.reshape(1, ph * pw).astype(np.float64)
    k = float(ph * pw)
    inv_area = 1.0 / k
    cc = a1 @ a2.T
    s1 = a1.sum(1); q1 = (a1 * a1).sum(1); s2 = a2.sum(1); q2 = (a2 * a2).sum(1)
    templ_mean = s1 * inv_area
    num = cc - templ_mean[:, None] * s2[None, :]
    if method == TM_CCOEFF_NORMED:
        templ_sdv = np.sqrt(np.maximum(q1 * inv_area - templ_mean * templ_mean, 0.0))
        templ_norm2 = templ_sdv * templ_sdv
        if templ_norm2[0] < _DBL_EPSILON:
            res = np.ones_like(num)
        else:
            templ_norm = np.sqrt(templ_norm2) / np.sqrt(inv_area)
            diff2 = np.maximum(q2 - s2 * s2 * inv_area, 0.0)
            t_w = np.where(diff2 <= np.minimum(0.5, 10.0 * _FLT_EPSILON * q2), 0.0, np.sqrt(diff2))
            t = templ_norm[:, None] * t_w[None, :]
            absn = np.abs(num)
            with np.errstate(divide='ignore', invalid='ignore'):
                res = np.where(absn < t, num / t, np.where(absn < t * 1.125, np.sign(num), 0.0))
    else:
        res = num
    res = res.astype(np.float32).reshape(oh, ow)
    return min_max(res.reshape(1, -1)).reshape(oh, ow)


def initial_co_map(img, template, ws, method=TM_CCOEFF_NORMED):
    """misc/Correlation_map.py:69-87 -- 4-D (T0,T1,T0,T1) float64 holding float32 values,
    each [i,j] slice independently min-maxed (misc/Feature_value.py:42)."""
    raw = match_template_matrix(img, template, ws, method)          # float32 [P,Q]
    t0 = img.shape[0] - ws + 1
    t1 = img.shape[1] - ws + 1
    norm = min_max(raw)                                              # float32 arithmetic
    return norm.astype(np.float64).reshape(t0, t1, t0, t1)


# --------------------------------------------------------------------------------------
# pyramid
# --------------------------------------------------------------------------------------
def rectify(m, lam=LAM):
    """misc/Correlation_map.py:158-159."""
    return m ** lam


def maxpool_3s2p1(m):
    """torch.nn.MaxPool2d(3, 2, padding=1) over the last two axes
    (misc/Correlation_map.py:176-184, used :100-103).  -inf padding, NaN propagates."""
    c, d = m.shape[-2:]
    oc, od = (c - 1) // 2 + 1, (d - 1) // 2 + 1
    pad = np.full(m.shape[:-2] + (c + 2, d + 2), -np.inf, dtype=m.dtype)
    pad[..., 1:c + 1, 1:d + 1] = m
    out = None
    nan = None
    for dy in range(3):
        for dx in range(3):
            v = pad[..., dy:dy + 2 * oc:2, dx:dx + 2 * od:2][..., :oc, :od]
            nan = np.isnan(v) if nan is None else (nan | np.isnan(v))
            out = v.copy() if out is None else np.fmax(out, v)
    if nan.any():
        out[nan] = np.nan
    return out


def aggregate(m):
    """misc/Correlation_map.py:89-130 -- max-pool every slice, then average the four
    children of each parent without any shift; add order ((ul+ur)+ll)+lr, then /4."""
    res = maxpool_3s2p1(m)
    l1, l2 = m.shape[2] // 2, m.shape[3] // 2
    ul = res[0:2 * l1:2, 0:2 * l2:2]
    ur = res[0:2 * l1:2, 1:2 * l2:2]
    ll = res[1:2 * l1:2, 0:2 * l2:2]
    lr = res[1:2 * l1:2, 1:2 * l2:2]
    return (ul + ur + ll + lr) / 4


def pyramid(co_map, lam=LAM):
    """misc/Correlation_map.py:132-156 -> (co_map_list, iteration, N_map)."""
    lst = []
    cur = rectify(co_map, lam)
    lst.append(cur)
    n = 1
    iteration = 1
    while n < min(co_map.shape[:2]):
        cur = rectify(aggregate(cur), lam)
        lst.append(cur)
        n *= 2
        iteration += 1
    return lst, iteration, n


def correlation_map(img, template, ws, method=TM_CCOEFF_NORMED):
    """misc/Correlation_map.py:161-173 -> dict(co_map, co_map_list, iteration, N_map)."""
    co = initial_co_map(img, template, ws, method)
    lst, it, n = pyramid(co)
    return dict(co_map=co, co_map_list=lst, iteration=it, N_map=n)


# --------------------------------------------------------------------------------------
# backtracking
# --------------------------------------------------------------------------------------
def _near_match_vec(level, p0, p1, d0, d1):
    """misc/Matching.py:58-78 vectorised over arrays of children.

    level: (A,B,C,D); p*: patch coordinates; d*: p_dot.  Returns (row, col, score) with
    first-max row-major tie-break over the zero-padded 3x3 window, NaN winning like
    np.argmax, and the ``max < 1e-4 -> centre`` rule.  Arithmetic in level.dtype.
    """
    a, b, c, d = level.shape
    pad = np.zeros((a, b, c + 2, d + 2), dtype=level.dtype)
    pad[:, :, 1:c + 1, 1:d + 1] = level
    wins = np.empty(p0.shape + (9,), dtype=level.dtype)
    n = 0
    for dy in range(3):
        for dx in range(3):
            wins[..., n] = pad[p0, p1, d0 + dy, d1 + dx]
            n += 1
    m = np.argmax(wins, axis=-1)                     # first max; first NaN if any
    mx = np.max(wins, axis=-1)                       # NaN propagates
    with np.errstate(invalid='ignore'):
        small = mx < level.dtype.type(NEAR_ZERO)
    m = np.where(small, 4, m)
    m0, m1 = m // 3, m % 3
    centre = pad[p0, p1, d0 + 1, d1 + 1]
    best = np.take_along_axis(wins, m[..., None], axis=-1)[..., 0]
    score = best + centre
    return d0 + m0 - 1, d1 + m1 - 1, score


def initial_move_map(top):
    """misc/Matching.py:80-96 -- (3,a,b) float64 at the pyramid top."""
    a, b = top.shape[:2]
    ii, jj = np.meshgrid(np.arange(a), np.arange(b), indexing='ij')
    r, c, s = _near_match_vec(top, ii, jj, ii, jj)
    out = np.zeros((3, a, b))
    out[0], out[1], out[2] = r, c, s
    return out


def backtrack_level(level, parent_map):
    """misc/Matching.py:98-139 (filtering off) -- (3,h,w) -> (3,2h,2w)."""
    h, w = parent_map.shape[1:]
    ii, jj = np.meshgrid(np.arange(2 * h), np.arange(2 * w), indexing='ij')
    pd0 = (parent_map[0] * 2).astype('int64')
    pd1 = (parent_map[1] * 2).astype('int64')
    d0 = pd0[ii // 2, jj // 2] + (ii & 1)
    d1 = pd1[ii // 2, jj // 2] + (jj & 1)
    r, c, s = _near_match_vec(level, ii, jj, d0, d1)
    out = np.empty((3, 2 * h, 2 * w))
    out[0], out[1], out[2] = r, c, s
    return out


def sub_pix(l0, mp):
    """misc/Matching.py:165-209 -- parabola fit on level 0; arithmetic in l0.dtype, result
    float64.  numpy index rules decide what happens at the edges: a negative index wraps
    (match row 0 reads its "-1" neighbour from the last row), an index >= the axis length
    raises IndexError, which the bare ``except`` swallows -> no refinement along that axis.
    Matches are inside the map unless the displacement filter ran on level 0, which can move
    them one step outside (then row/col -1 wraps and row/col == size skips both axes)."""
    t0, t1, c, d = l0.shape
    out = mp.copy()
    ii, jj = np.meshgrid(np.arange(t0), np.arange(t1), indexing='ij')
    c0 = mp[0].astype('int64')
    c1 = mp[1].astype('int64')
    two = l0.dtype.type(2)

    def fit(r0, r1, r_):
        with np.errstate(divide='ignore', invalid='ignore'):
            ok = (r0 > r1) & (r0 > r_)
            diff = -(r1 - r_) / (two * (r1 + r_ - two * r0))
        return np.where(ok, diff, 0).astype(np.float64)

    def ok_idx(v, n):                                # numpy accepts -n <= v < n
        return (v >= -n) & (v < n)

    base = ok_idx(c0, c) & ok_idx(c1, d)
    r0 = l0[ii, jj, c0 % c, c1 % d]
    # rows
    valid = base & ok_idx(c0 + 1, c) & ok_idx(c0 - 1, c)
    r1 = l0[ii, jj, (c0 + 1) % c, c1 % d]
    r_ = l0[ii, jj, (c0 - 1) % c, c1 % d]
    d_x = ii - mp[0]
    out[0] = np.where(valid, ii - d_x + fit(r0, r1, r_), ii - d_x)
    # cols
    valid = base & ok_idx(c1 + 1, d) & ok_idx(c1 - 1, d)
    r1 = l0[ii, jj, c0 % c, (c1 + 1) % d]
    r_ = l0[ii, jj, c0 % c, (c1 - 1) % d]
    d_y = jj - mp[1]
    out[1] = np.where(valid, jj - d_y + fit(r0, r1, r_), jj - d_y)
    return out


def matching_margins(co_map_list):
    """matching(co_map_list, sub_pix_on=False) plus, per level-0 patch, the smallest margin by which
    any decision on its top-down path was taken: at every level the difference between the best
    and the best different value of the zero-padded 3x3 window (misc/Matching.py:58-78), and the
    distance of the window maximum from the 1e-4 threshold.  A float32 pyramid may decide a step
    differently from the reference's float64 one only where this margin is of the size of the
    float32 error of the level values -- the "provable near-tie" the parity tests allow.
    Returns (map (3,T0,T1), margin (T0,T1))."""
    def margins(level, p0, p1, d0, d1):
        a, b, c, d = level.shape
        pad = np.zeros((a, b, c + 2, d + 2), dtype=level.dtype)
        pad[:, :, 1:c + 1, 1:d + 1] = level
        wins = np.stack([pad[p0, p1, d0 + dy, d1 + dx] for dy in range(3) for dx in range(3)], -1)
        top = wins.max(-1, keepdims=True)
        gap = top - wins
        # exact ties are structural (overlapping pooling windows select the same element: the same
        # expression on the same inputs, equal in any precision) and are broken identically by the
        # first-maximum rule; the margin is the distance to the best DIFFERENT value
        gap = np.where(gap > 0, gap, np.inf).min(-1)
        return np.minimum(gap, np.abs(top[..., 0] - NEAR_ZERO))

    top = co_map_list[-1]
    a, b = top.shape[:2]
    ii, jj = np.meshgrid(np.arange(a), np.arange(b), indexing='ij')
    mp = initial_move_map(top)
    mg = margins(top, ii, jj, ii, jj)
    idx = len(co_map_list) - 1
    while idx > 0:
        idx -= 1
        level = co_map_list[idx]
        h, w = mp.shape[1:]
        ii, jj = np.meshgrid(np.arange(2 * h), np.arange(2 * w), indexing='ij')
        d0 = (mp[0] * 2).astype('int64')[ii // 2, jj // 2] + (ii & 1)
        d1 = (mp[1] * 2).astype('int64')[ii // 2, jj // 2] + (jj & 1)
        mg = np.minimum(mg[ii // 2, jj // 2], margins(level, ii, jj, d0, d1))
        mp = backtrack_level(level, mp)
    return mp, mg


def match_filter(mp, window=3, mode='median'):
    """misc/Matching.py:224-255 -- outlier filter on the displacement field of a (3,h,w) map:
    interior cells get round(mean | median of the (2e+1)^2 neighbourhood of displacements)
    + their own coordinate (Python round = half to even; the neighbourhood is read from a
    snapshot, so the update order does not matter); border cells and the score plane are
    untouched; maps smaller than the window pass through.  The reference sizes its snapshot
    (shape[1], shape[1]) and is therefore only defined on square maps."""
    mp = mp.copy()
    _, h, w = mp.shape
    if not (h >= window and w >= window):
        return mp
    if h != w:
        raise ValueError('Matching._filter is undefined on non-square maps (%d x %d)' % (h, w))
    e = int((window - 1) / 2)
    k = 2 * e + 1
    d_col = (mp[1] - np.arange(w)[None, :]).astype('int64')
    d_row = (mp[0] - np.arange(h)[:, None]).astype('int64')
    red = {'average': np.mean, 'median': np.median}[mode]
    from numpy.lib.stride_tricks import sliding_window_view
    if h - 2 * e > 0 and w - 2 * e > 0:
        ii, jj = np.meshgrid(np.arange(e, h - e), np.arange(e, w - e), indexing='ij')
        mp[1, e:h - e, e:w - e] = np.rint(red(sliding_window_view(d_col, (k, k)), axis=(-2, -1))) + jj
        mp[0, e:h - e, e:w - e] = np.rint(red(sliding_window_view(d_row, (k, k)), axis=(-2, -1))) + ii
    return mp


def matching(co_map_list, sub_pix_on=True, return_levels=False, filtering=False, filtering_num=3,
             filter_window_size=3, filtering_mode='median'):
    """misc/Matching.py:211-222 -> (3,T0,T1) float64.  With ``filtering`` the first
    ``filtering_num`` maps of the top-down pass (the top map, then the result of each _B)
    go through match_filter; the counter runs down even where the map is too small
    (misc/Matching.py:91-93,136-138)."""
    left = filtering_num if filtering else 0

    def maybe_filter(m):
        nonlocal left
        if left > 0:
            m = match_filter(m, filter_window_size, filtering_mode)
            left -= 1
        return m

    mp = maybe_filter(initial_move_map(co_map_list[-1]))
    levels = [mp]
    idx = len(co_map_list) - 1
    while idx > 0:
        idx -= 1
        mp = maybe_filter(backtrack_level(co_map_list[idx], mp))
        levels.append(mp)
    pre = mp
    if sub_pix_on:
        mp = sub_pix(co_map_list[0], mp)
    if return_levels:
        return mp, pre, levels
    return mp


# --------------------------------------------------------------------------------------
# disparity planes, post-hoc sub-pixel, tiling
# --------------------------------------------------------------------------------------
MODES = ['elevation', 'elevation2', 'distance']      # misc/Calc_difference.py:30


def cal_map(mp, mode='elevation'):
    """misc/Calc_difference.py:25-49."""
    if mode not in MODES:
        raise SystemExit
    t0, t1 = mp.shape[1:]
    ii, jj = np.meshgrid(np.arange(t0, dtype=np.float64), np.arange(t1, dtype=np.float64), indexing='ij')
    if mode == 'elevation':
        return jj - mp[1]
    if mode == 'elevation2':
        return ii - mp[0]
    d0 = ii - mp[0]
    d1 = jj - mp[1]
    # np.linalg.norm of a 1-D 2-vector evaluates sqrt(dot(x, x)); the BLAS ddot kernel
    # of this image fuses the second product: sqrt(fma(d1, d1, d0*d0))  (checked against
    # the live reference on 1e5 random vectors: 0 mismatches; plain d0*d0+d1*d1 differs
    # by 1 ulp in 8 % of them).
    return np.sqrt(_fma(d1, d1, d0 * d0))


def _fma(a, b, c):
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library('m') or 'libm.so.6')
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    f = np.frompyfunc(lambda x, y, z: libm.fma(float(x), float(y), float(z)), 3, 1)
    return f(a, b, c).astype(np.float64)


def image_threshold(arr, threshold=(0, 10)):
    """misc/optimize_loop.py:40-44."""
    arr = np.where(arr > threshold[1], threshold[1], arr)
    arr = np.where(arr < threshold[0], threshold[0], arr)
    return arr


def sub_pix_cal(arr, co_map, direction=0, ratio=100.):
    """misc/sub_pix_cal.py:22-53 -- post-hoc 2-D parabola refinement of a mosaic."""
    arr = image_threshold(np.asarray(arr), threshold=[-3, 3]).astype(float)
    co = np.asarray(co_map, dtype=np.float64)
    s0, s1 = arr.shape
    if s0 > 2 and s1 > 2:
        d = arr[1:-1, 1:-1]
        r0 = co[1:-1, 1:-1] * ratio
        if direction == 0:
            r1 = co[2:, 1:-1] * ratio
            r_ = co[:-2, 1:-1] * ratio
        else:
            r1 = co[1:-1, 2:] * ratio
            r_ = co[1:-1, :-2] * ratio
        with np.errstate(divide='ignore', invalid='ignore'):
            dis = d - (r1 - r_) / (2 * (r1 + r_ - 2 * r0))
            dis = np.where(np.abs(d - dis) > 1, d, dis)
        arr = arr.copy()
        arr[1:-1, 1:-1] = dis
    return image_threshold(arr, threshold=[-3, 3])


def solve_tile(img1, img2, ws, modes=('elevation',), sub_pix_on=True, method=TM_CCOEFF_NORMED, filtering=None):
    """misc/image_cut_solver.py:115-142 -> ((nmodes,T0,T1), (T0,T1)).
    filtering = None or (filtering_num, filter_window_size, filtering_mode)."""
    cm = correlation_map(img1, img2, ws, method)
    if filtering is None:
        out = matching(cm['co_map_list'], sub_pix_on)
    else:
        out = matching(cm['co_map_list'], sub_pix_on, filtering=True, filtering_num=filtering[0],
                       filter_window_size=filtering[1], filtering_mode=filtering[2])
    return np.array([cal_map(out, m) for m in modes]), out[2]


def tile_grid(img_shape, image_size, stride, ws):
    """misc/image_cut_solver.py:53-62 -> (len0, len1); the last fitting tile is dropped."""
    e = int((ws - 1) / 2)
    trimmed = [image_size[i] + 2 * e for i in range(2)]
    return [int(np.floor((img_shape[i] - trimmed[i]) / stride[i])) for i in range(2)], trimmed


def image_cut_solver(img1, img2, image_size=(32, 32), stride=(32, 32), ws=5,
                     modes=('elevation',), sub_pix_on=True, method=TM_CCOEFF_NORMED,
                     tile_rows=None, filtering=None, tiles=None):
    """misc/image_cut_solver.py:95-113,144-184 -> (d_map (nmodes,S0',S1'), out_map (S0',S1')).

    ``tile_rows=(lo,hi)`` restricts the solve to tile-row indices lo..hi-1, ``tiles=(a,b)`` to the
    tiles a..b-1 in row-major order (everything else is left NaN; pixels of the range that later tiles of
    the range paste over are overwritten as usual) -- used to check the multi-GPU partitions.
    """
    ln, trimmed = tile_grid(img1.shape, image_size, stride, ws)
    if ln[0] <= 0 or ln[1] <= 0:
        raise IndexError('list index out of range')         # img_index[-1] on an empty list
    size = [stride[i] * (ln[i] - 1) + image_size[i] for i in range(2)]
    d_map = np.full([len(modes)] + size, np.nan)
    out_map = np.full(size, np.nan)
    lo, hi = (0, ln[0]) if tile_rows is None else tile_rows
    for j in range(ln[1]):
        for i in range(lo, hi):
            if tiles is not None and not (tiles[0] <= i * ln[1] + j < tiles[1]):
                continue                                         # a range of tiles in row-major order (the multi-GPU shares)
            y, x = stride[0] * i, stride[1] * j
            d, s = solve_tile(img1[y:y + trimmed[0], x:x + trimmed[1]],
                              img2[y:y + trimmed[0], x:x + trimmed[1]], ws, modes, sub_pix_on, method, filtering)
            d_map[:, y:y + image_size[0], x:x + image_size[1]] = d
            out_map[y:y + image_size[0], x:x + image_size[1]] = s
    return d_map, out_map


def bilateral_filter_u8(src, d, sigma_color, sigma_space):
    """cv2.bilateralFilter on a uint8 plane (optimize_looper.py:76-77), restating OpenCV's own
    8-bit algorithm (bilateral_filter.dispatch.cpp / .simd.hpp, non-IPP path): float32 colour
    and space weights, BORDER_REFLECT_101, neighbours accumulated row by row, cvRound."""
    import math
    src = np.ascontiguousarray(src, dtype=np.uint8)
    if sigma_color <= 0:
        sigma_color = 1
    if sigma_space <= 0:
        sigma_space = 1
    radius = int(np.rint(sigma_space * 1.5)) if d <= 0 else d // 2
    radius = max(radius, 1)
    gc = np.float32(-0.5 / (sigma_color * sigma_color))
    gs = np.float32(-0.5 / (sigma_space * sigma_space))
    cw = np.array([np.float32(math.exp(float(np.float32(i * i) * gc))) for i in range(256)], dtype=np.float32)
    h, w = src.shape
    t = np.pad(src, radius, mode='reflect').astype(np.int64)
    c = t[radius:radius + h, radius:radius + w]
    s = np.zeros((h, w), np.float32)
    ws = np.zeros((h, w), np.float32)
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            r = math.sqrt(float(i * i + j * j))
            if r > radius:
                continue
            w0 = np.float32(math.exp(r * r * float(gs)))
            v = t[radius + i:radius + i + h, radius + j:radius + j + w]
            wk = (w0 * cw[np.abs(v - c)]).astype(np.float32)
            s = (s + (v.astype(np.float32) * wk).astype(np.float32)).astype(np.float32)
            ws = (ws + wk).astype(np.float32)
    return np.rint(s / ws).astype(np.uint8)


def raw_read(path, size=(6000, 6000), rate=1):
    """misc/raw_read.py:36-45 -- int8 read, * rate, cast to uint8 (wraps)."""
    c = np.fromfile(path, dtype=np.int8, count=size[0] * size[1]).reshape(1, size[1], size[0])
    return (c * rate)[0].astype(np.uint8)


# --------------------------------------------------------------------------------------
# Gauss-Seidel post-process (the reference's dead `if 0:` branch, optimize_looper.py:55-74)
# --------------------------------------------------------------------------------------
def optimize_loop(img_dis, coefficient, alpha, exclusion, size):
    """misc/optimize_loop.py:15-37 -- clamp to [0, 10], one forward and one backward in-place sweep of
    d <- (-a d + alpha (left + right + up + down)) / (-a + 4 alpha), a = coefficient[i, j].
    The reference visits the cells one by one in row-major (then reversed) order; a cell only reads
    its four neighbours, so all cells of an anti-diagonal i + j = t are independent and the sweep can
    run diagonal by diagonal -- the order the CUDA kernel uses.  Bit-identical to the sequential loops.
    Returns (img_dis, error) with error = the sequentially accumulated |old - new| of the backward sweep."""
    d = image_threshold(np.asarray(img_dis), threshold=[0, 10]).astype(np.float64)
    co = np.asarray(coefficient, dtype=np.float64)
    lo0, hi0, lo1, hi1 = exclusion, size[0] - exclusion - 1, exclusion, size[1] - exclusion - 1     # half-open
    if hi0 <= lo0 or hi1 <= lo1:
        return d, 0.0

    def update(i, j):
        sum_d = d[i, j - 1] + d[i, j + 1] + d[i - 1, j] + d[i + 1, j]
        a = co[i, j]
        return (-a * d[i, j] + alpha * sum_d) / (-a + 4.0 * alpha)

    for t in range(lo0 + lo1, hi0 + hi1 - 1):
        i = np.arange(max(lo0, t - (hi1 - 1)), min(hi0 - 1, t - lo1) + 1)
        d[i, t - i] = update(i, t - i)
    # "Reverse" sweep (misc/optimize_loop.py:26-36).  The reference flips the indices INSIDE the inner loop:
    # j is rebound by `for j`, but i keeps its value across inner iterations, so `i = size[0] - i - 1`
    # flips it on the first inner iteration, flips it BACK on the second, and so on.  Visit k of outer
    # iteration o therefore is the cell (R - o, C - k) for even k and (lo0 + o, C - k) for odd k
    # (R = size0 - lo0 - 1, C = size1 - lo1 - 1): even-k columns are walked bottom-up, odd-k columns
    # top-down.  A cell's vertical neighbours belong to its own column chain; its horizontal neighbours
    # belong to the other kind of column and were visited at o' = R - lo0 - o -- strictly earlier or later
    # except at the one step 2 o = R - lo0, where all columns sit on the same row and the visit order along
    # k decides.  So all columns advance in lockstep over o (independent within a step), and that one
    # middle row is walked sequentially.
    R, C = size[0] - lo0 - 1, size[1] - lo1 - 1
    n0, n1 = hi0 - lo0, hi1 - lo1
    k = np.arange(n1)
    cols = C - k
    diff = np.zeros((n0, n1))
    for o in range(n0):
        rows = np.where(k % 2 == 0, R - o, lo0 + o)
        if 2 * o == R - lo0:
            for kk in range(n1):
                new = update(rows[kk], cols[kk])
                diff[o, kk] = np.abs(d[rows[kk], cols[kk]] - new)
                d[rows[kk], cols[kk]] = new
        else:
            new = update(rows, cols)
            diff[o] = np.abs(d[rows, cols] - new)
            d[rows, cols] = new
    visit = diff.ravel()                                             # the reference's visiting order
    error = float(np.add.accumulate(visit)[-1]) if visit.size else 0.0
    return d, error


def make_weight(guide_img, exclusion, size, sigma):
    """misc/opt_loop.py:66-85 -> (gausian_weight (w, w), color_weight_matrix (S0 - e, S1 - e, w, w))."""
    e = exclusion
    w = 2 * e + 1
    g = np.asarray(guide_img, dtype=np.float64)
    k = np.arange(w) - e
    gw = np.exp(-((k[:, None] ** 2 + k[None, :] ** 2).astype(np.float64)) / (2.0 * sigma[1] ** 2))
    cw = np.zeros([size[0] - e, size[1] - e, w, w])
    from numpy.lib.stride_tricks import sliding_window_view
    win = sliding_window_view(g, (w, w))                             # win[i - e, j - e] = g[i-e:i+e+1, j-e:j+e+1]
    n0, n1 = size[0] - 2 * e - 1, size[1] - 2 * e - 1                # cells i in [e, S0 - e - 1)
    if n0 > 0 and n1 > 0:
        sub = win[:n0, :n1]
        c = sub[:, :, e, e][:, :, None, None] - sub
        cw[:n0, :n1] = np.exp(-1 * c * c / (2.0 * sigma[0] ** 2))
    return gw, cw


def optimize_loop_bilateral(img_dis, color_weight_matrix, gausian_weight, coefficient, alpha, exclusion, size, vertical=False):
    """misc/opt_loop.py:16-63 (horizon / vertical): in-place sweep of the bilateral-weighted update over
    the (2e+1)^2 window.  A cell reads every cell of its window, the earlier ones already updated: all
    cells with the same j + (e + 1) i are independent (the last updated cell a cell needs, (i-1, j+e), lies
    one step earlier), which is the order used here and by the CUDA kernel.  `a` and `b` use the
    CONSTANT entries coefficient[e, e], coefficient[e, e +- 1] (or [e +- 1, e]) exactly as the reference does.
    Returns (img_dis, error)."""
    e = exclusion
    d = np.asarray(img_dis, dtype=np.float64)                        # in place, like the reference
    co = np.asarray(coefficient, dtype=np.float64)
    gw = np.asarray(gausian_weight, dtype=np.float64)
    lo0, hi0, lo1, hi1 = e, size[0] - e - 1, e, size[1] - e - 1
    cp, cm = (co[e + 1, e], co[e - 1, e]) if vertical else (co[e, e + 1], co[e, e - 1])
    a = -(co[e, e] - (cp + cm) / 2.0)
    shift = (cp - cm) / 2.0 / (-2.0 * co[e, e] + cp + cm)
    diff = np.zeros_like(d)
    off = np.arange(-e, e + 1)
    for t in range(lo1 + (e + 1) * lo0, (hi1 - 1) + (e + 1) * (hi0 - 1) + 1):
        i = np.arange(lo0, hi0)
        j = t - (e + 1) * i
        ok = (j >= lo1) & (j < hi1)
        i, j = i[ok], j[ok]
        if i.size == 0:
            continue
        sub = d[(i[:, None] + off[None, :])[:, :, None], (j[:, None] + off[None, :])[:, None, :]]     # (n, w, w)
        cw = color_weight_matrix[i - e, j - e]
        b = d[i, j] - shift
        num = np.array([(gw * cw[k] * sub[k]).sum() for k in range(i.size)])
        den = np.array([(gw * cw[k]).sum() for k in range(i.size)])
        new = (-a * b + num) / (-a + den)
        diff[i, j] = np.abs(d[i, j] - new)
        d[i, j] = new
    visit = diff[lo0:hi0, lo1:hi1].ravel()
    error = float(np.add.accumulate(visit)[-1]) if visit.size else 0.0
    return d, error
