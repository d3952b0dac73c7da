"""Helpers of the GPU parity tests: compare a mosaic with the oracle tile by tile and classify
every disagreeing pixel as a provable near-tie (or not)."""
import numpy as np

from oracle import dm_oracle as O

# A float32 pyramid may take a top-down decision differently from the float64 reference only
# where the two best candidates of a 3x3 window are closer than the float32 error of the level
# values (measured <= 1e-4 relative).
NEAR_TIE = 2e-4


def owned(gi, gj, len0, len1, size, stride):
    """Pixels of tile (gi, gj) no later tile pastes over (misc/image_cut_solver.py:165-175)."""
    o0 = size[0] if gi == len0 - 1 else min(size[0], stride[0])
    o1 = size[1] if gj == len1 - 1 else min(size[1], stride[1])
    return o0, o1


def scene_report(img1, img2, size, stride, ws, modes, d, sc, sub_pix=True, tiles=None, method=O.TM_CCOEFF_NORMED):
    """d (n_modes, S0', S1'), sc (S0', S1') from the CUDA path against the oracle on every tile (or
    the listed global tile indices).  Only 'elevation' / 'elevation2' planes take part in the
    integer comparison.  Returns a dict of counts; `bad_not_near_tie` must be zero."""
    (len0, len1), trimmed = O.tile_grid(img1.shape, size, stride, ws)
    todo = range(len0 * len1) if tiles is None else tiles
    rep = dict(n=0, int_bad=0, bad_not_near_tie=0, sub_bad=0, max_sub_rel=0.0, score_bad=0, max_score=0.0, worst_margin_of_bad=0.0)
    for g in todo:
        gi, gj = divmod(int(g), len1)
        y, x = stride[0] * gi, stride[1] * gj
        a = img1[y:y + trimmed[0], x:x + trimmed[1]]
        b = img2[y:y + trimmed[0], x:x + trimmed[1]]
        cm = O.correlation_map(a, b, ws, method)
        pre, margin = O.matching_margins(cm['co_map_list'])
        out = O.sub_pix(cm['co_map_list'][0], pre) if sub_pix else pre
        o0, o1 = owned(gi, gj, len0, len1, size, stride)
        bad_any = np.zeros((o0, o1), bool)
        for m, mode in enumerate(modes):
            ref = O.cal_map(out, mode)[:o0, :o1]
            got = d[m, y:y + o0, x:x + o1]
            ok = ~(np.isnan(ref) | np.isnan(got))
            assert np.array_equal(np.isnan(ref), np.isnan(got))
            bad = ok & (np.abs(got - ref) > 0.5)
            bad_any |= bad
            rel = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
            good = ok & ~bad
            rep['n'] += int(ok.sum())
            rep['int_bad'] += int(bad.sum())
            rep['sub_bad'] += int((rel[good] > 1e-3).sum())
            if good.any():
                rep['max_sub_rel'] = max(rep['max_sub_rel'], float(rel[good].max()))
        mg = margin[:o0, :o1]
        rep['bad_not_near_tie'] += int((bad_any & ~(mg < NEAR_TIE)).sum())
        if bad_any.any():
            rep['worst_margin_of_bad'] = max(rep['worst_margin_of_bad'], float(mg[bad_any].max()))
        rs = out[2][:o0, :o1]
        gs = sc[y:y + o0, x:x + o1]
        sok = ~(np.isnan(rs) | np.isnan(gs)) & ~bad_any
        sd = np.abs(gs - rs)
        rep['score_bad'] += int((sd[sok] > 1e-3).sum())
        if sok.any():
            rep['max_score'] = max(rep['max_score'], float(sd[sok].max()))
    rep['int_disagreement'] = rep['int_bad'] / max(1, rep['n'])
    rep['sub_frac'] = rep['sub_bad'] / max(1, rep['n'])
    return rep


def assert_parity(rep, max_int=1e-3, max_sub=1e-3):
    """north star: <= 0.1 % integer-disparity disagreement, every one of them a near-tie of the
    reference's own decision; sub-pixel disparity and scores within 1e-3."""
    print('parity report: %s' % rep)
    assert rep['int_disagreement'] <= max_int, rep
    assert rep['bad_not_near_tie'] == 0, rep
    assert rep['sub_frac'] <= max_sub, rep
    assert rep['score_bad'] <= max_sub * max(1, rep['n']), rep
