# Drop-in shim: the reference's scripts do `from misc.Correlation_map import ...`
# (deep_dem_mathing.py:11-13, ex_deepmatching_rawinput.py:17-18); the implementation lives
# in deepmatching_stereo_matching_b200.Correlation_map and runs on the GPU through libdmstereo.
from deepmatching_stereo_matching_b200.Correlation_map import Correlation_map, Maxpool, Feature_value  # noqa: F401
