# Drop-in shim: the reference's scripts do `from misc.optimize_loop import ...`
# (deep_dem_mathing.py:11-13, ex_deepmatching_rawinput.py:17-18); the implementation lives
# in deepmatching_stereo_matching_b200.optimize_loop and runs on the GPU through libdmstereo.
from deepmatching_stereo_matching_b200.optimize_loop import image_threshold, optimize_loop  # noqa: F401
