"""Supplementary single-GPU measurements of the other BASELINE.json configs (device-resident,
CUDA events; the headline bench line is bench.py).  Prints one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepmatching_stereo_matching_b200 import _native
from deepmatching_stereo_matching_b200.synth import stereo_pair

CONFIGS = {
    'c1': dict(shape=(256, 256), n=1, T=32, s=32, ws=5, modes=['elevation'], sub=True),
    'c2': dict(shape=(1024, 1024), n=1, T=64, s=60, ws=15, modes=['elevation', 'elevation2'], sub=True),
    'c3': dict(shape=(4096, 4096), n=1, T=64, s=60, ws=15, modes=['elevation', 'elevation2'], sub=True),
    'c4': dict(shape=(512, 512), n=64, T=32, s=32, ws=5, modes=['elevation', 'elevation2'], sub=True, post=True),
    'c5_strip': dict(shape=(1134, 8192), n=1, T=128, s=124, ws=15, modes=['elevation', 'elevation2'], sub=True),   # one of the 8 strips of the 8192^2 scene (8 tile rows)
}

def run(name, steps=3):
    c = CONFIGS[name]
    h, w = c['shape']
    if c['n'] == 1:
        i1, i2 = stereo_pair((h, w), seed=2, mode='sine', amp=c['T'] // 4)
        i1, i2 = i1[None], i2[None]
    else:
        base = [stereo_pair((h, w), seed=100 + b, mode='sine', amp=c['T'] // 4) for b in range(4)]
        i1 = np.stack([base[b % 4][0] for b in range(c['n'])]); i2 = np.stack([base[b % 4][1] for b in range(c['n'])])
    d1 = torch.from_numpy(np.ascontiguousarray(i1)).cuda(); d2 = torch.from_numpy(np.ascontiguousarray(i2)).cuda()
    prm = _native.scene_params((h, w), (c['T'], c['T']), (c['s'], c['s']), c['ws'], 'cv2.TM_CCOEFF_NORMED', c['modes'], c['sub'], None, -1, n_scenes=c['n'])
    info = _native.scene_geometry(prm)
    ctx = _native.Context(timing=True)
    nm = len(c['modes'])
    dm = torch.zeros((c['n'], nm, info.out_h, info.out_w), dtype=torch.float64, device='cuda')
    om = torch.zeros((c['n'], info.out_h, info.out_w), dtype=torch.float64, device='cuda')
    post = torch.empty_like(dm) if c.get('post') else None
    lib = _native.lib()
    def step():
        inf = ctx.solve_device(prm, d1, d2, dm, om)
        if post is not None:        # config 4: sub_pix_cal(elevation, score, 1), sub_pix_cal(elevation2, score, 0)
            for b in range(c['n']):
                for m, name_ in enumerate(c['modes']):
                    _native.check(lib.dm_sub_pix_cal(_native.ptr(dm[b, m]), _native.ptr(om[b]), info.out_h, info.out_w,
                                                     1 if name_ == 'elevation' else 0, 100.0, _native.ptr(post[b, m]), _native.stream_ptr()))
        return inf
    for _ in range(2): inf = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): inf = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st, _ = ctx.stage_ms()
    px = c['n'] * info.out_h * info.out_w
    print(json.dumps({'config': name, 'scene': [c['n'], h, w], 'T': c['T'], 'ws': c['ws'], 'tiles': inf.n_tiles, 'chunk_tiles': inf.chunk_tiles,
                      'fused': bool(inf.used_fused), 'output_px': px, 'ms': round(ms, 3), 'MP_per_s': round(px / ms / 1e3, 2),
                      'stage_ms': {k: round(v, 3) for k, v in st.items()}, 'workspace_GB': round(ctx.workspace_bytes / 1e9, 2)}))
    ctx.close()

if __name__ == '__main__':
    for name in (sys.argv[1:] or list(CONFIGS)):
        run(name)
