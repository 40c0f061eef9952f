"""Python binding of the C++17 host class (include/mccba_host.hpp), same method names as the reference's
MultiCameraCalibration: loadImages / initialize / optimizeExtrinsics / run / writeParameters."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmccba_host.so")
_lib = None

PINHOLE, OMNIDIRECTIONAL = 0, 1


def lib():
    global _lib
    if _lib is None:
        capi.lib()      # libmccba.so first (RTLD_GLOBAL) so the host library resolves against it
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m multi_camera_calibration_b200.build`" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.mccbah_last_error.restype = C.c_char_p
        _lib.mccbah_last_error.argtypes = [C.c_void_p]
    return _lib


class MultiCameraCalibration:
    def __init__(self, cameraType, nCameras, fileName, patternWidth, patternHeight, verbose=0, showExtration=0,
                 nMiniMatches=20, flags=0, criteria=(1, 20, 1e-7), mode=capi.MODE_REFERENCE_GN, device=0):
        self._h = C.c_void_p()
        rc = lib().mccbah_create(int(cameraType), int(nCameras), str(fileName).encode(), C.c_float(patternWidth),
                                 C.c_float(patternHeight), int(verbose), int(showExtration), int(nMiniMatches), int(flags),
                                 int(criteria[0]), int(criteria[1]), C.c_double(criteria[2]), int(mode), int(device),
                                 C.byref(self._h))
        self._check(rc)

    def _check(self, rc):
        if rc:
            raise RuntimeError(lib().mccbah_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().mccbah_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def loadImages(self):
        self._check(lib().mccbah_load_images(self._h))

    def reset(self):
        self._check(lib().mccbah_reset(self._h))

    def initialize(self):
        self._check(lib().mccbah_initialize(self._h))

    def optimizeExtrinsics(self):
        e = C.c_double()
        self._check(lib().mccbah_optimize_extrinsics(self._h, C.byref(e)))
        return e.value

    def run(self):
        e = C.c_double()
        self._check(lib().mccbah_run(self._h, C.byref(e)))
        return e.value

    def removeOutlier(self, threshold=0.5):
        n = C.c_int()
        self._check(lib().mccbah_remove_outlier(self._h, C.c_float(threshold), C.byref(n)))
        return n.value

    def writeParameters(self, filename):
        self._check(lib().mccbah_write_parameters(self._h, str(filename).encode()))

    def indexing(self):
        nv, ne = C.c_int(), C.c_int()
        self._check(lib().mccbah_sizes(self._h, C.byref(nv), C.byref(ne)))
        ec = np.zeros(ne.value, dtype=np.int32); ep = np.zeros(ne.value, dtype=np.int32); ei = np.zeros(ne.value, dtype=np.int32)
        vt = np.zeros(nv.value, dtype=np.int32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self._check(lib().mccbah_get_indexing(self._h, ip(ec), ip(ep), ip(ei), ip(vt)))
        return dict(edge_cam=ec, edge_pv=ep, photo_index=ei, vertex_timestamp=vt)

    def parameters(self):
        nv, ne = C.c_int(), C.c_int()
        self._check(lib().mccbah_sizes(self._h, C.byref(nv), C.byref(ne)))
        p = np.zeros(6 * (nv.value - 1))
        self._check(lib().mccbah_get_parameters(self._h, p.ctypes.data_as(C.POINTER(C.c_double))))
        return p

    def initialParameters(self):
        nv, ne = C.c_int(), C.c_int()
        self._check(lib().mccbah_sizes(self._h, C.byref(nv), C.byref(ne)))
        p = np.zeros(6 * (nv.value - 1))
        self._check(lib().mccbah_get_initial_parameters(self._h, p.ctypes.data_as(C.POINTER(C.c_double))))
        return p

    def stats(self):
        me, rms, it, ms = C.c_double(), C.c_double(), C.c_int(), C.c_double()
        self._check(lib().mccbah_get_stats(self._h, C.byref(me), C.byref(rms), C.byref(it), C.byref(ms)))
        return dict(mean_reproj_error=me.value, rms=rms.value, iterations=it.value, device_ms=ms.value)


class MyMultiCameraCalibration(MultiCameraCalibration):
    """The subclass the reference's sample runs (include/opencv2/ccalib/mymulticalib.hpp:95-100): corner files
    <dataFolder>/<serial>/<timestamp>.yaml, intrinsics <cameraConfigFolder>/<serial>.xml, solvePnP initial poses."""

    def __init__(self, cameraSerials, cameraType, nCameras, dataFolder, cameraConfigFolder, doubleSideConfig="",
                 frontPatternSize=(9, 6), backPatternSize=(0, 0), patternWidth=360.0, patternHeight=200.0, verbose=0,
                 criteria=(3, 200, 1e-7), mode=capi.MODE_REFERENCE_GN, device=0):
        self._h = C.c_void_p()
        rc = lib().mccbah_create_my(",".join(cameraSerials).encode(), int(cameraType), int(nCameras), str(dataFolder).encode(),
                                    str(cameraConfigFolder).encode(), str(doubleSideConfig).encode(), int(frontPatternSize[0]),
                                    int(frontPatternSize[1]), int(backPatternSize[0]), int(backPatternSize[1]),
                                    C.c_float(patternWidth), C.c_float(patternHeight), int(verbose), int(criteria[0]),
                                    int(criteria[1]), C.c_double(criteria[2]), int(mode), int(device), C.byref(self._h))
        self._check(rc)

    def loadImages(self, outliers=()):
        self._check(lib().mccbah_load_images_my(self._h, "\n".join(outliers).encode()))

    def removeOutlier(self):
        """Drops the edges whose mean reprojection error exceeds 0.5 px; returns the set of their corner files."""
        buf = C.create_string_buffer(1 << 22)
        n = C.c_int()
        self._check(lib().mccbah_remove_outlier_my(self._h, buf, len(buf), C.byref(n)))
        return set(s for s in buf.value.decode().split("\n") if s)

    def run(self):
        e = C.c_double()
        self._check(lib().mccbah_run_my(self._h, C.byref(e)))
        return e.value


def solve_pnp(obj, img, K5, dist8, ndist):
    """cv::solvePnP (SOLVEPNP_ITERATIVE) as restated in csrc/host/pnp.cpp.  Returns (rvec, tvec)."""
    obj = np.ascontiguousarray(obj, dtype=np.float64).reshape(-1, 3)
    img = np.ascontiguousarray(img, dtype=np.float64).reshape(-1, 2)
    K5 = np.ascontiguousarray(K5, dtype=np.float64); d8 = np.zeros(8); d8[:len(dist8)] = dist8
    r = np.zeros(3); t = np.zeros(3)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib().mccbah_solve_pnp(int(obj.shape[0]), dp(obj), dp(img), dp(K5), dp(d8), int(ndist), dp(r), dp(t))
    if rc:
        raise RuntimeError("solve_pnp failed (%d)" % rc)
    return r, t


class DoubleSideCalibration(MyMultiCameraCalibration):
    """Double-sided board in front of fixed cameras (include/opencv2/ccalib/doubleSide.hpp:82-178): camera poses from the
    "CameraMatrix" of <cameraConfigFolder>/<serial>.xml, images of both pattern sides, unknown front<->back transform +
    one board pose per timestamp, optimised on the GPU (include/mccba.h: mccba_ds_*)."""

    def __init__(self, cameraSerials, cameraType, nCameras, dataFolder, cameraConfigFolder, frontPatternSize=(9, 6),
                 backPatternSize=(8, 5), patternWidth=360.0, patternHeight=200.0, verbose=0, criteria=(3, 200, 1e-8), device=0):
        self._h = C.c_void_p()
        rc = lib().mccbah_create_ds(",".join(cameraSerials).encode(), int(cameraType), int(nCameras), str(dataFolder).encode(),
                                    str(cameraConfigFolder).encode(), int(frontPatternSize[0]), int(frontPatternSize[1]),
                                    int(backPatternSize[0]), int(backPatternSize[1]), C.c_float(patternWidth),
                                    C.c_float(patternHeight), int(verbose), int(criteria[0]), int(criteria[1]),
                                    C.c_double(criteria[2]), int(device), C.byref(self._h))
        self._check(rc)

    def initialize(self):
        self._check(lib().mccbah_initialize_ds(self._h))

    def optimizeExtrinsics(self):
        """Returns the RMS reprojection error (pixels) at the optimum."""
        r = C.c_double()
        self._check(lib().mccbah_optimize_ds(self._h, C.byref(r)))
        return r.value

    def run(self):
        self.loadImages()
        self.initialize()
        return self.optimizeExtrinsics()

    def doubleSideTransform(self):
        """4 x 4: back-pattern coordinates -> front-pattern coordinates."""
        T = np.zeros((4, 4))
        self._check(lib().mccbah_get_double_side_transform(self._h, T.ctypes.data_as(C.POINTER(C.c_double))))
        return T
