"""Times the correlation engines alone on C2-shaped batches (CUDA events, current stream)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.synth import texture

lib = _native.lib()
t0 = t1 = 64; ws = 15; n = int(sys.argv[1]) if len(sys.argv) > 1 else 225
P, kpad = t0 * t1, lib.dm_kpad(ws)
H = W = 1024
s1 = torch.from_numpy(texture((H, W), seed=1)).cuda(); s2 = torch.from_numpy(texture((H, W), seed=2)).cuda()
origin = torch.tensor([[60 * (k // 15), 60 * (k % 15)] for k in range(n)], dtype=torch.int32, device='cuda')
bufs = []
for side, sc in ((1, s1), (2, s2)):
    desc = torch.empty((n * P, kpad), dtype=torch.bfloat16, device='cuda')
    stat = torch.empty((n * P * 6,), dtype=torch.float32, device='cuda')
    _native.check(lib.dm_descriptors(_native.ptr(sc), H, W, W, _native.ptr(origin), n, t0, t1, ws, side, _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
    bufs += [desc, stat]
raw = torch.empty((n, P, P), dtype=torch.float32, device='cuda')
flops = 2.0 * ws * ws * P * P * n
for name, engine in (('umma_raw', 2), ('umma_null', 3), ('pool', 4)):
    if os.environ.get('DM_ONLY_POOL'): break
    for _ in range(3):
        _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, engine, _native.ptr(raw), _native.stream_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, engine, _native.ptr(raw), _native.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('%-10s %.3f ms  %.0f TFLOP/s' % (name, ms, flops / ms / 1e9))

# pooled epilogue through the scene solver's stage timer
import numpy as np
from deepmatching_stereo_matching_b200.synth import stereo_pair
i1, i2 = stereo_pair((1024, 1024), seed=1, mode='sine', amp=16)
d1 = torch.from_numpy(i1).cuda(); d2 = torch.from_numpy(i2).cuda()
prm = _native.scene_params((1024, 1024), (64, 64), (60, 60), 15, 'cv2.TM_CCOEFF_NORMED', ['elevation', 'elevation2'], True, None, 1)
ctx = _native.Context(timing=True)
planes = torch.zeros((3, 904, 904), dtype=torch.float64, device='cuda')
for _ in range(3):
    ctx.solve_device(prm, d1, d2, planes[:-1], planes[-1])
acc = {}
for _ in range(5):
    ctx.solve_device(prm, d1, d2, planes[:-1], planes[-1])
    ms, _ = ctx.stage_ms()
    for k, v in ms.items(): acc[k] = acc.get(k, 0) + v / 5
print('fused stages', {k: round(v, 3) for k, v in acc.items()})
