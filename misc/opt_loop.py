# Drop-in shim: optimize_looper.py:23 does `from misc.opt_loop import *`; the implementation lives in
# deepmatching_stereo_matching_b200.opt_loop and runs on the GPU through libdmstereo.
from deepmatching_stereo_matching_b200.opt_loop import (  # noqa: F401
    make_weight, optimize_loop_bilateral_horizon, optimize_loop_bilateral_vertical)
