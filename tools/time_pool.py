"""Times the tcgen05 correlation kernel alone (CUDA events, current stream): pooled epilogue (engine 4),
MMA + TMEM drain only (3), for a tile shape.   python tools/time_pool.py T ws n_tiles"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.synth import texture

lib = _native.lib()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ws = int(sys.argv[2]) if len(sys.argv) > 2 else 15
n = int(sys.argv[3]) if len(sys.argv) > 3 else 225
P, kpad = T * T, lib.dm_kpad(ws)
H = W = 2048
s1 = torch.from_numpy(texture((H, W), seed=1)).cuda(); s2 = torch.from_numpy(texture((H, W), seed=2)).cuda()
per_row = (W - T - ws) // 60
origin = torch.tensor([[60 * ((k // per_row) % per_row), 60 * (k % per_row)] for k in range(n)], dtype=torch.int32, device='cuda')
bufs = []
for side, sc in ((1, s1), (2, s2)):
    desc = torch.empty((n * P, kpad), dtype=torch.bfloat16, device='cuda')
    stat = torch.empty((n * P * 6,), dtype=torch.float32, device='cuda')
    _native.check(lib.dm_descriptors(_native.ptr(sc), H, W, W, _native.ptr(origin), n, T, T, ws, side, _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
    bufs += [desc, stat]
raw = torch.empty((n * P * (P // 4) + 8 * n * P,), dtype=torch.float32, device='cuda')
flops = 2.0 * ws * ws * P * P * n
for name, engine in (('pool', 4), ('drain-only', 3)):
    for _ in range(3):
        _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, engine, _native.ptr(raw), _native.stream_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, engine, _native.ptr(raw), _native.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('T=%d ws=%d n=%d %-10s %.3f ms  %.0f TFLOP/s  pooled write %.0f GB/s' % (T, ws, n, name, ms, flops / ms / 1e9, 4.0 * n * P * P / 4 / ms / 1e6))
