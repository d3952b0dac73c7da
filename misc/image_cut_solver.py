# Drop-in shim: the reference's scripts do `from misc.image_cut_solver import ...`
# (deep_dem_mathing.py:11-13, ex_deepmatching_rawinput.py:17-18); the implementation lives
# in deepmatching_stereo_matching_b200.image_cut_solver and runs on the GPU through libdmstereo.
from deepmatching_stereo_matching_b200.image_cut_solver import ImageCutSolver, Correlation_map, Matching, Calc_difference  # noqa: F401
