"""B200-native calibration bundle adjustment (hot path of yulong314/multi_camera_calibration).

The product is the CUDA library libmccba.so behind the C ABI in include/mccba.h; this package is the thin Python
plumbing over it (ctypes) plus the host-side mirror of the reference's MultiCameraCalibration interface."""
from . import capi  # noqa: F401
from .capi import Solver, MccbaError  # noqa: F401

__all__ = ["capi", "Solver", "MccbaError"]
