import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.synth import stereo_pair
i1, i2 = stereo_pair((1024, 1024), seed=1, mode='sine', amp=16)
d1 = torch.from_numpy(i1).cuda(); d2 = torch.from_numpy(i2).cuda()
prm = _native.scene_params((1024, 1024), (64, 64), (60, 60), 15, 'cv2.TM_CCOEFF_NORMED', ['elevation', 'elevation2'], True, None, 1)
ctx = _native.Context()
planes = torch.zeros((3, 904, 904), dtype=torch.float64, device='cuda')
for _ in range(2):
    ctx.solve_device(prm, d1, d2, planes[:-1], planes[-1])
torch.cuda.synchronize()
