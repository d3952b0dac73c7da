"""Per-stage device time of one config (stage events of the library), e.g. to compare a kernel variant
selected by an environment switch:  python tools/time_stages.py c2|c3|c4"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deepmatching_stereo_matching_b200 import _native
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
c = bench.CONFIGS[name]
i1, i2 = bench.make_scene(name)
d1, d2 = torch.from_numpy(np.ascontiguousarray(i1)).cuda(), torch.from_numpy(np.ascontiguousarray(i2)).cuda()
prm = _native.scene_params(c['shape'], [c['T']] * 2, [c['stride']] * 2, c['ws'], bench.FEATURE, bench.MODES, True, n_scenes=c['batch'])
info = _native.scene_geometry(prm)
ctx = _native.Context(timing=True)
dm_ = torch.zeros((c['batch'], 2, info.out_h, info.out_w), dtype=torch.float64, device='cuda')
om_ = torch.zeros((c['batch'], info.out_h, info.out_w), dtype=torch.float64, device='cuda')
for _ in range(3):
    ctx.solve_device(prm, d1, d2, dm_, om_)
acc = {}
n = 10
for _ in range(n):
    ctx.solve_device(prm, d1, d2, dm_, om_)
    ms, _ = ctx.stage_ms()
    for k, v in ms.items():
        acc[k] = acc.get(k, 0) + v / n
print(name, 'DM_FINAL_QUAD=%s' % os.environ.get('DM_FINAL_QUAD', '-'), {k: round(v, 4) for k, v in acc.items()}, 'sum %.3f' % sum(acc.values()))
