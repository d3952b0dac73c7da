# Drop-in shim: the reference's scripts do `from misc.sub_pix_cal import ...`
# (deep_dem_mathing.py:11-13, ex_deepmatching_rawinput.py:17-18); the implementation lives
# in deepmatching_stereo_matching_b200.sub_pix_cal and runs on the GPU through libdmstereo.
from deepmatching_stereo_matching_b200.sub_pix_cal import sub_pix_cal, image_threshold  # noqa: F401
