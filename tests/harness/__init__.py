"""Builds and loads the host-side math harness (tests/harness/math_harness.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmath_harness.so")
_SRC = os.path.join(_HERE, "math_harness.cpp")
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "multi_camera_calibration_b200", "csrc")
_HDR = os.path.join(_CSRC, "mccba_math.cuh")
_HDR2 = os.path.join(_CSRC, "mccba_f32x2.cuh")
_lib = None


def lib():
    global _lib
    if _lib is None:
        newest = max(os.path.getmtime(_SRC), os.path.getmtime(_HDR), os.path.getmtime(_HDR2))
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _SO, _SRC])
        _lib = C.CDLL(_SO)
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def rig_step(rig, params, lam, policy=0):
    """Per-edge blocks (E x 28), reduced system (S, gs) and the unscaled additive step, computed on the host with the
    product's own __host__ __device__ arithmetic (tangent-space formulation).  policy 1: the per-edge blocks come from
    the packed single-precision pass (mccba_f32x2.cuh) in the device kernel's lane order."""
    E = rig["edge_cam"].size
    ns = 6 * (rig["n_cam"] - 1)
    p = np.ascontiguousarray(params, dtype=np.float64)
    blocks = np.zeros((E, 28)); S = np.zeros((max(ns, 1), max(ns, 1))); gs = np.zeros(max(ns, 1)); step = np.zeros(p.size)
    S = np.zeros((ns, ns)) if ns else np.zeros((1, 1))
    bad = lib().hm_rig_step(int(rig["n_cam"]), int(rig["n_frame"]), int(E), _p(np.ascontiguousarray(rig["edge_cam"], dtype=np.int32), C.c_int),
                            _p(np.ascontiguousarray(rig["edge_pv"], dtype=np.int32), C.c_int),
                            _p(np.ascontiguousarray(rig["edge_off"], dtype=np.int64), C.c_int64),
                            _p(np.ascontiguousarray(rig["obj"], dtype=np.float32), C.c_float),
                            _p(np.ascontiguousarray(rig["img"], dtype=np.float32), C.c_float),
                            _p(np.ascontiguousarray(rig["cam_model"], dtype=np.int32), C.c_int),
                            _p(np.ascontiguousarray(rig["cam_K5"], dtype=np.float64)),
                            _p(np.ascontiguousarray(rig["cam_dist8"], dtype=np.float64)),
                            _p(np.ascontiguousarray(rig["cam_ndist"], dtype=np.int32), C.c_int),
                            _p(np.ascontiguousarray(rig["cam_xi"], dtype=np.float64)), _p(p), C.c_double(lam), _p(blocks),
                            _p(S), _p(gs), _p(step), int(policy))
    return dict(bad=bad, blocks=blocks, S=S[:ns, :ns], gs=gs[:ns], step=step)


def rodrigues(om):
    R = np.zeros(9)
    lib().hm_rodrigues(_p(np.ascontiguousarray(om, dtype=np.float64)), _p(R))
    return R.reshape(3, 3)


def left_jacobian_inv_apply(om, psi):
    out = np.zeros(3)
    lib().hm_left_jacobian_inv_apply(_p(np.ascontiguousarray(om, dtype=np.float64)),
                                     _p(np.ascontiguousarray(psi, dtype=np.float64)), _p(out))
    return out


def omni_project(obj, om, T, K5, xi, D4):
    obj = np.ascontiguousarray(obj, dtype=np.float64).reshape(-1, 3)
    n = obj.shape[0]
    proj = np.zeros((n, 2)); jac = np.zeros((2 * n, 16))
    lib().hm_omni_project(n, _p(obj), _p(np.ascontiguousarray(om, dtype=np.float64)), _p(np.ascontiguousarray(T, dtype=np.float64)),
                          _p(np.ascontiguousarray(K5, dtype=np.float64)), C.c_double(float(xi)),
                          _p(np.ascontiguousarray(D4, dtype=np.float64)), _p(proj), _p(jac))
    return proj, jac
