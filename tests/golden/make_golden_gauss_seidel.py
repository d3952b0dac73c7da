# -*- coding: utf-8 -*-
"""
Golden vectors for the Gauss-Seidel post-process of the reference (its dead `if 0:` branch in
optimize_looper.py:55-74): misc/optimize_loop.py::optimize_loop (4-neighbour) and
misc/opt_loop.py::make_weight / optimize_loop_bilateral_horizon / _vertical, run UNMODIFIED from
/root/reference.  Build container only:

    python tests/golden/make_golden_gauss_seidel.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, '/root/reference')
from misc.optimize_loop import optimize_loop                                   # noqa: E402
from misc.opt_loop import make_weight, optimize_loop_bilateral_horizon, optimize_loop_bilateral_vertical   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(11)
out = {}
# a disparity plane and a correlation-score mosaic of the size a small scene produces
size = (37, 45)
d = rng.normal(size=size) * 2.0 + 4.0
d[8:20, 10:30] += 3.0
co = rng.random(size) * 1.2 + 0.4
out['d'] = d
out['co'] = co
for k, (alpha, exclusion, loops) in enumerate([(0.008, 3, 1), (0.008, 3, 3), (0.05, 1, 2)]):
    x = d.copy()
    errs = []
    for _ in range(loops):
        x, err = optimize_loop(x, co, alpha, exclusion, size)
        errs.append(err)
    out['ol%d_cfg' % k] = np.array([alpha, exclusion, loops], dtype=np.float64)
    out['ol%d_out' % k] = x
    out['ol%d_err' % k] = np.array(errs)
for k, (sigma, exclusion, loops) in enumerate([((5, 5), 3, 2), ((2, 3), 2, 1)]):
    sg = np.array([int(s) for s in sigma])                                     # optimize_looper.py:52
    gw, cw = make_weight(d, exclusion, size, sg)
    out['bl%d_cfg' % k] = np.array([sigma[0], sigma[1], exclusion, loops], dtype=np.float64)
    out['bl%d_gw' % k] = gw
    out['bl%d_cw' % k] = cw
    for name, fn in (('h', optimize_loop_bilateral_horizon), ('v', optimize_loop_bilateral_vertical)):
        x = d.copy()
        errs = []
        for _ in range(loops):
            x, err = fn(x, cw, gw, co, 0.008, exclusion, size)
            errs.append(err)
        out['bl%d_%s_out' % (k, name)] = x
        out['bl%d_%s_err' % (k, name)] = np.array(errs)
path = os.path.join(HERE, 'gauss_seidel.npz')
np.savez_compressed(path, **out)
print(path, os.path.getsize(path))
