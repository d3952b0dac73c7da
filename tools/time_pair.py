"""Times the pooled / drain-only tcgen05 correlation with and without CTA pairs (C2-shaped batch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.synth import texture

lib = _native.lib()
t0 = t1 = int(os.environ.get('DM_T', 64)); ws = int(os.environ.get('DM_WS', 15)); n = int(sys.argv[1]) if len(sys.argv) > 1 else 225
P, kpad = t0 * t1, lib.dm_kpad(ws)
H = W = max(1024, t0 + ws + 60 * 15)
s1 = torch.from_numpy(texture((H, W), seed=1)).cuda(); s2 = torch.from_numpy(texture((H, W), seed=2)).cuda()
origin = torch.tensor([[60 * ((k % 225) // 15), 60 * (k % 15)] for k in range(n)], dtype=torch.int32, device='cuda')
bufs = []
for side, sc in ((1, s1), (2, s2)):
    desc = torch.empty((n * P, kpad), dtype=torch.bfloat16, device='cuda')
    stat = torch.empty((n * P * 6,), dtype=torch.float32, device='cuda')
    _native.check(lib.dm_descriptors(_native.ptr(sc), H, W, W, _native.ptr(origin), n, t0, t1, ws, side, _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
    bufs += [desc, stat]
raw = torch.empty((n, P, P // 2), dtype=torch.float32, device='cuda')
flops = 2.0 * ws * ws * P * P * n
for pair in (0, -1):
    _native.check(lib.dm_correlation_set_pair_mode(pair))
    for name, engine in (('null', 3), ('pool', 4)):
        for _ in range(3):
            _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, engine, _native.ptr(raw), _native.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, engine, _native.ptr(raw), _native.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print('pair=%2d %-5s %.3f ms  %.0f TFLOP/s' % (pair, name, ms, flops / ms / 1e9))
