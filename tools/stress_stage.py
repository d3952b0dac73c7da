"""Bitwise repeatability of the descriptor kernel and the correlation engines (same inputs, many launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.synth import texture
lib = _native.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for (T, ws, n) in [(32, 5, 20), (32, 7, 6), (64, 15, 9), (16, 5, 36), (64, 5, 7)]:
    e2 = ws - 1
    H, W = T + e2 + 9, (T + e2) + 11 * (n - 1)
    s1 = torch.from_numpy(texture((H, W), seed=31)).cuda(); s2 = torch.from_numpy(texture((H, W), seed=32, plain_noise=True)).cuda()
    origin = torch.tensor([[k % 9, 11 * k] for k in range(n)], dtype=torch.int32, device='cuda')
    P, kpad = T * T, lib.dm_kpad(ws)
    def descs():
        out = []
        for side, sc in ((1, s1), (2, s2)):
            desc = torch.full((n * P, kpad), 7.0, dtype=torch.bfloat16, device='cuda')
            stat = torch.full((n * P * 6,), 7.0, dtype=torch.float32, device='cuda')
            _native.check(lib.dm_descriptors(_native.ptr(sc), H, W, W, _native.ptr(origin), n, T, T, ws, side, _native.ptr(desc), _native.ptr(stat), _native.stream_ptr()))
            out += [desc, stat]
        return out
    b0 = descs()
    nd = 0
    for _ in range(5):
        b = descs()
        nd += int(any(not torch.equal(x.view(torch.int16) if x.dtype == torch.bfloat16 else x.view(torch.int32), y.view(torch.int16) if y.dtype == torch.bfloat16 else y.view(torch.int32)) for x, y in zip(b, b0)))
    res = {}
    for name, engine, size in (('simt', 1, n * P * P), ('umma_raw', 2, n * P * P), ('umma_pool', 4, n * P * (P // 4) + 4 * n * P)):
        first, bad, nbadel = None, 0, 0
        for r in range(reps if engine != 1 else 2):
            raw = torch.full((size,), float('nan'), dtype=torch.float32, device='cuda')
            _native.check(lib.dm_correlation(_native.ptr(b0[0]), _native.ptr(b0[1]), _native.ptr(b0[2]), _native.ptr(b0[3]), n, P, kpad, ws, 5, engine, _native.ptr(raw), _native.stream_ptr()))
            torch.cuda.synchronize()
            cur = raw.view(torch.int32)
            if first is None: first = cur.clone()
            else:
                ne = int((cur != first).sum())
                if ne: bad += 1; nbadel = max(nbadel, ne)
        res[name] = (bad, nbadel, first)
    eq = bool(torch.equal(res['simt'][2], res['umma_raw'][2]))
    nan_pool = int(torch.isnan(res['umma_pool'][2].view(torch.float32)).sum())
    print('T=%d ws=%d n=%d  desc_nondet=%d  raw: %d/%d runs differ (max %d elems)  pool: %d/%d runs differ (max %d elems)  simt==raw:%s  pool NaNs:%d'
          % (T, ws, n, nd, res['umma_raw'][0], reps - 1, res['umma_raw'][1], res['umma_pool'][0], reps - 1, res['umma_pool'][1], eq, nan_pool))
