# -*- coding: utf-8 -*-
"""
deepmatching_stereo_matching_b200 -- the DeepMatching-for-stereo hot path of
Yuki-Kumon/deepmatching_stereo_matching on B200 (sm_100a): hand-written CUDA behind a
C-ABI library (csrc/, include/dmstereo.h) and the reference's Python classes on top.
Importing the package does not touch the GPU; using any class does, and raises if the
library or a CUDA device is missing (there is no CPU fallback).
"""

from .Feature_value import Feature_value            # noqa: F401
from .Correlation_map import Correlation_map        # noqa: F401
from .Matching import Matching                      # noqa: F401
from .Calc_difference import Calc_difference        # noqa: F401
from .sub_pix_cal import sub_pix_cal                # noqa: F401
from .optimize_loop import image_threshold, optimize_loop   # noqa: F401
from .opt_loop import make_weight, optimize_loop_bilateral_horizon, optimize_loop_bilateral_vertical   # noqa: F401
from .image_cut_solver import ImageCutSolver        # noqa: F401
from .raw_read import RawRead                       # noqa: F401
from .bilateral import bilateral_filter             # noqa: F401

__version__ = '0.1.0'
