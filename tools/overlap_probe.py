"""Feasibility probe: does an HBM-bound kernel stream run concurrently with the persistent tcgen05 correlation
kernel when that kernel leaves some SMs free (DM_CORR_MAX_PAIRS)?  Stream A: pooled correlation of n tiles;
stream B: aggregation level transitions (HBM-bound) on other buffers.  Prints A alone, B alone, both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.synth import texture
lib = _native.lib()
T, ws, n = 64, 15, 225
P, kpad = T * T, lib.dm_kpad(ws)
H = W = 1024
s1 = torch.from_numpy(texture((H, W), seed=1)).cuda(); s2 = torch.from_numpy(texture((H, W), seed=2)).cuda()
origin = torch.tensor([[60 * (k // 15), 60 * (k % 15)] for k in range(n)], dtype=torch.int32, device='cuda')
bufs = []
for side, sc in ((1, s1), (2, s2)):
    desc = torch.empty((n * P, kpad), dtype=torch.bfloat16, device='cuda')
    stat = torch.empty((n * P * 6,), dtype=torch.float32, device='cuda')
    _native.check(lib.dm_descriptors(_native.ptr(sc), H, W, W, _native.ptr(origin), n, T, T, ws, side, _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
    bufs += [desc, stat]
raw = torch.empty((n * P * (P // 4) + 8 * n * P,), dtype=torch.float32, device='cuda')
# HBM-bound work: level 1 -> 2 of 225 tiles (0.94 GB read + 0.06 GB write per call) on its own buffers
lvl = torch.rand((n, 32, 32, 32, 32), dtype=torch.float32, device='cuda')
out = torch.empty((n, 16, 16, 16, 16), dtype=torch.float32, device='cuda')
sa, sb = torch.cuda.Stream(priority=-1), torch.cuda.Stream()
def run_a(reps):
    with torch.cuda.stream(sa):
        for _ in range(reps):
            _native.check(lib.dm_correlation(*[_native.ptr(b) for b in bufs], n, P, kpad, ws, 5, 4, _native.ptr(raw), _native.stream_ptr()))
def run_b(reps):
    with torch.cuda.stream(sb):
        for _ in range(reps):
            _native.check(lib.dm_aggregate(_native.ptr(lvl), n, 32, 32, 32, 32, 1, _native.ptr(out), _native.stream_ptr()))
def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn()
    sa.synchronize(); sb.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
run_a(2); run_b(2)
RA, RB = 5, int(os.environ.get('RB', 40))
ta = timed(lambda: run_a(RA)); tb = timed(lambda: run_b(RB))
tab = timed(lambda: (run_a(RA), run_b(RB)))
tba = timed(lambda: (run_b(RB), run_a(RA)))
print('DM_CORR_MAX_PAIRS=%s: A (%d correlations) %.2f ms, B (%d aggregations, %.1f GB) %.2f ms, A then B queued %.2f ms, B then A queued %.2f ms, sum %.2f' % (
    os.environ.get('DM_CORR_MAX_PAIRS', '-'), RA, ta, RB, RB * 1.0, tb, tab, tba, ta + tb))
