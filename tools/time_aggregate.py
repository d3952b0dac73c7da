import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmatching_stereo_matching_b200 import _native
lib = _native.lib()
# level transitions of the fused path: C2 (225 tiles, image_size 64): level 1 -> 2 and 2 -> 3; C4 (a quarter of the
# 12544 tiles of 64 x 512^2, image_size 32): level 1 -> 2.  Knobs: DM_AGG_NO_MERGE, DM_AGG_BUDGET_KB, DM_PDL.
for (n, A, C) in ((225, 32, 32), (225, 16, 16), (3136, 16, 16)):
    x = torch.rand((n, A, A, C, C), dtype=torch.float32, device='cuda')
    y = torch.empty((n, A // 2, A // 2, C // 2, C // 2), dtype=torch.float32, device='cuda')
    for _ in range(3):
        _native.check(lib.dm_aggregate(_native.ptr(x), n, A, A, C, C, 1, _native.ptr(y), _native.stream_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # flush L2 between launches with a big memset
    big = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    tot = 0.0
    for _ in range(10):
        big.zero_()
        e0.record()
        _native.check(lib.dm_aggregate(_native.ptr(x), n, A, A, C, C, 1, _native.ptr(y), _native.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 10
    gb = (x.numel() + y.numel()) * 4 / 1e9
    print('n %d level %dx%d  %.4f ms  %.0f GB/s' % (n, A, C, ms, gb / ms * 1e3), flush=True)
