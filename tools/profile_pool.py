"""Launches the pooled correlation kernels once each on a C2-shaped batch (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.synth import texture

lib = _native.lib()
t0 = t1 = int(os.environ.get('DM_T', 64)); ws = int(os.environ.get('DM_WS', 15)); n = int(sys.argv[1]) if len(sys.argv) > 1 else 225
engines = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [4]
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
raw = torch.empty((n * P * (P // 4) + 8 * n * P,), dtype=torch.float32, device='cuda')
for engine in engines:
    for _ in range(2):
        _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, engine, _native.ptr(raw), _native.stream_ptr()))
torch.cuda.synchronize()
print('done')
