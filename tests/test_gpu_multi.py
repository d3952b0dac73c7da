"""Several devices in ONE process (dm_multi_*, ImageCutSolver.devices): the tile rows of a scene
(or the pairs of a batch) spread over the visible devices must give the bits of the one-device
solve.  Needs at least two GPUs; skipped on a one-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dm():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    if torch.cuda.device_count() < 2:
        pytest.skip('needs at least two GPUs')
    import deepmatching_stereo_matching_b200 as pkg
    return pkg


def _scene(shape, seed=7, amp=16):
    from deepmatching_stereo_matching_b200.synth import stereo_pair
    return stereo_pair(shape, seed=seed, mode='sine', amp=amp)


def test_image_cut_solver_uses_all_devices_bit_identical(dm):
    """The ex_deepmatching_rawinput.py call sequence (ws 15, image_size 64, stride 60) on a large
    scene: one Python process, every visible GPU, unchanged signatures -- same bits as one GPU."""
    import torch
    from deepmatching_stereo_matching_b200 import image_cut_solver as ics
    n = torch.cuda.device_count()
    side = 4096 if n >= 4 else 2048
    i1, i2 = _scene((side, side))
    kw = dict(image_size=[64, 64], stride=[60, 60], window_size=15, degree_map_mode=['elevation', 'elevation2'], sub_pix=True)
    s = dm.ImageCutSolver(i1, i2, **kw)
    s.log_flg = False
    d, sc = s()                                      # devices = None: all visible devices
    assert ics.visible_devices() == list(range(n))
    assert tuple(range(n)) in ics._MULTI, 'the solve did not go through dm_multi'
    assert s.info.n_tiles == s.len[0] * s.len[1] and s.info.used_fused == 1
    one = dm.ImageCutSolver(i1, i2, **kw)
    one.log_flg = False
    one.devices = [0]
    d1, sc1 = one()
    assert np.array_equal(d, d1, equal_nan=True) and np.array_equal(sc, sc1, equal_nan=True)


@pytest.mark.parametrize('fused,filtering', [(0, False), (1, True), (1, False)])
def test_explicit_device_lists_and_paths(dm, fused, filtering):
    """Explicit device lists (also a list that does not start at device 0), the materialising path and
    the displacement filter through dm_multi_solve_scene_host."""
    import torch
    n = torch.cuda.device_count()
    i1, i2 = _scene((400, 520), seed=11, amp=5)
    kw = dict(image_size=[32, 32], stride=[30, 30], window_size=5, degree_map_mode=['elevation', 'distance'], sub_pix=True,
              filtering=filtering, filtering_window_size=3, filtering_num=3, filtering_mode='median')
    res = []
    for devs in ([0], list(range(n)), [n - 1, 0]):
        s = dm.ImageCutSolver(i1, i2, **kw)
        s.log_flg = False
        s.fused = fused
        s.devices = devs
        d, sc = s()
        assert s.info.used_fused == fused
        res.append((d.copy(), sc.copy()))
    for d, sc in res[1:]:
        assert np.array_equal(d, res[0][0], equal_nan=True) and np.array_equal(sc, res[0][1], equal_nan=True)


def test_batch_of_pairs_over_devices(dm):
    """Config-4 style batch: whole pairs per device."""
    from deepmatching_stereo_matching_b200 import image_cut_solver as ics
    pairs = [_scene((200, 232), seed=100 + b, amp=6) for b in range(5)]
    i1 = np.stack([p[0] for p in pairs]); i2 = np.stack([p[1] for p in pairs])
    kw = dict(image_size=[32, 32], stride=[32, 32], window_size=5, degree_map_mode=['elevation', 'elevation2'], sub_pix=True)
    d0, s0 = ics.solve_batch(i1, i2, devices=[0], **kw)
    import torch
    d1, s1 = ics.solve_batch(i1, i2, devices=list(range(torch.cuda.device_count())), **kw)
    assert np.array_equal(d0, d1, equal_nan=True) and np.array_equal(s0, s1, equal_nan=True)


@pytest.mark.parametrize('gather', [0, 1], ids=['p2p_stream', 'nccl'])
def test_device_resident_solve_and_gather(dm, gather):
    """dm_multi_solve_scene: scenes and planes stay on the devices, the strips meet on the root --
    streamed over NVLink peer memory while the strip is being solved, or one grouped NCCL
    send/recv of the owned rows after it (ncclCommInitAll inside the library)."""
    import torch
    from deepmatching_stereo_matching_b200 import _native
    n = torch.cuda.device_count()
    i1, i2 = _scene((1024, 1024), seed=1)
    prm = _native.scene_params(i1.shape, [64, 64], [60, 60], 15, 'cv2.TM_CCOEFF_NORMED', ['elevation', 'elevation2'], True)
    info = _native.scene_geometry(prm)
    ref = torch.zeros((3, info.out_h, info.out_w), dtype=torch.float64, device='cuda:0')
    with torch.cuda.device(0):
        ctx = _native.Context()
        ctx.solve_device(prm, torch.from_numpy(i1).cuda(), torch.from_numpy(i2).cuda(), ref[:-1], ref[-1])
        torch.cuda.synchronize()
    multi = _native.MultiContext(list(range(n)))
    imgs1 = [torch.from_numpy(i1).to('cuda:%d' % r) for r in range(n)]
    imgs2 = [torch.from_numpy(i2).to('cuda:%d' % r) for r in range(n)]
    for root in (0, n - 1):
        planes = [torch.full((3, info.out_h, info.out_w), float('nan'), dtype=torch.float64, device='cuda:%d' % r) for r in range(n)]
        for r in range(n):
            torch.cuda.synchronize(r)
        got = multi.solve_device(prm, imgs1, imgs2, planes, root=root, gather=gather)
        multi.synchronize()
        assert got.n_tiles == info.n_tiles
        assert torch.equal(planes[root].cpu(), ref.cpu()), 'root %d gather %d' % (root, gather)
    multi.close()
    ctx.close()
