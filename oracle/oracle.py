"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/libmccba_oracle.so (oracle/mccba_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (multi_camera_calibration_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmccba_oracle.so")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_i64_p = C.POINTER(C.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "mccba_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmccba_oracle.so"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_rig_create.restype = C.c_void_p
        _lib.orc_rig_eval.restype = C.c_double
        _lib.orc_omni_build.restype = C.c_double
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=c_double_p):
    return a.ctypes.data_as(t)


def rodrigues(om, want_jac=True):
    om = _d(om).reshape(3)
    R = np.zeros(9)
    J = np.zeros(27)
    lib().orc_rodrigues(_p(om), _p(R), _p(J) if want_jac else None)
    return R.reshape(3, 3), J.reshape(3, 9)


def rodrigues_inv(R):
    R = _d(R).reshape(9)
    om = np.zeros(3)
    lib().orc_rodrigues_inv(_p(R), _p(om))
    return om


def compose_motion(om1, T1, om2, T2):
    om3 = np.zeros(3); T3 = np.zeros(3); d = np.zeros(72)
    lib().orc_compose_motion(_p(_d(om1).reshape(3)), _p(_d(T1).reshape(3)), _p(_d(om2).reshape(3)),
                             _p(_d(T2).reshape(3)), _p(om3), _p(T3), _p(d))
    return om3, T3, d.reshape(8, 3, 3)


def project_pinhole(obj, om, T, K5, dist, want_jac=True):
    obj = _d(obj).reshape(-1, 3)
    n = obj.shape[0]
    dist = _d(dist).reshape(-1)
    proj = np.zeros((n, 2)); jac = np.zeros((2 * n, 6))
    rc = lib().orc_project_pinhole(n, _p(obj), _p(_d(om).reshape(3)), _p(_d(T).reshape(3)), _p(_d(K5).reshape(5)),
                                   _p(dist) if dist.size else None, int(dist.size), _p(proj),
                                   _p(jac) if want_jac else None)
    if rc:
        raise ValueError("unsupported distortion length %d" % dist.size)
    return proj, jac


def project_omnidir(obj, om, T, K5, xi, D4, want_jac=True):
    obj = _d(obj).reshape(-1, 3)
    n = obj.shape[0]
    proj = np.zeros((n, 2)); jac = np.zeros((2 * n, 16))
    lib().orc_project_omnidir(n, _p(obj), _p(_d(om).reshape(3)), _p(_d(T).reshape(3)), _p(_d(K5).reshape(5)),
                              C.c_double(float(xi)), _p(_d(D4).reshape(4)), _p(proj), _p(jac) if want_jac else None)
    return proj, jac


class Rig:
    """Oracle-side rig problem.  Arrays follow the reference's members: per-edge cameraVertex / photoVertex
    (include/opencv2/ccalib/multicalib.hpp:86-103), float32 object/image points concatenated edge by edge."""

    def __init__(self, n_cam, n_frame, edge_cam, edge_pv, edge_off, obj_xyz, img_uv, cam_model, cam_K5, cam_dist8,
                 cam_ndist, cam_xi):
        self.n_cam, self.n_frame = int(n_cam), int(n_frame)
        self.edge_cam = np.ascontiguousarray(edge_cam, dtype=np.int32)
        self.edge_pv = np.ascontiguousarray(edge_pv, dtype=np.int32)
        self.edge_off = np.ascontiguousarray(edge_off, dtype=np.int64)
        self.n_edge = self.edge_cam.size
        self.obj = np.ascontiguousarray(obj_xyz, dtype=np.float32).reshape(-1, 3)
        self.img = np.ascontiguousarray(img_uv, dtype=np.float32).reshape(-1, 2)
        self.cam_model = np.ascontiguousarray(cam_model, dtype=np.int32)
        self.cam_K5 = _d(cam_K5).reshape(self.n_cam, 5)
        self.cam_dist8 = _d(cam_dist8).reshape(self.n_cam, 8)
        self.cam_ndist = np.ascontiguousarray(cam_ndist, dtype=np.int32)
        self.cam_xi = _d(cam_xi).reshape(self.n_cam)
        self.n_param = 6 * (self.n_cam + self.n_frame - 1)
        h = lib().orc_rig_create(self.n_cam, self.n_frame, self.n_edge, _p(self.edge_cam, c_int_p),
                                 _p(self.edge_pv, c_int_p), _p(self.edge_off, c_i64_p), _p(self.obj, c_float_p),
                                 _p(self.img, c_float_p), _p(self.cam_model, c_int_p), _p(self.cam_K5),
                                 _p(self.cam_dist8), _p(self.cam_ndist, c_int_p), _p(self.cam_xi))
        if not h:
            raise ValueError("orc_rig_create failed (bad vertex indices)")
        self._h = C.c_void_p(h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_rig_destroy(self._h)
            self._h = None

    def eval(self, params, policy=0, want_blocks=True):
        params = _d(params)
        assert params.size == self.n_param
        return float(lib().orc_rig_eval(self._h, _p(params), int(policy), int(want_blocks)))

    def blocks(self, which):
        ln = {0: 36, 1: 6, 2: 36, 3: 36, 4: 36, 5: 6, 6: 6, 7: 1, 8: 1}[which]
        out = np.zeros((self.n_edge, ln))
        lib().orc_rig_get_blocks(self._h, which, _p(out))
        return out

    def solve_normal(self, params, lam=0.0):
        params = _d(params)
        ns = 6 * (self.n_cam - 1)
        step = np.zeros(self.n_param); S = np.zeros((max(ns, 1), max(ns, 1))); gs = np.zeros(max(ns, 1))
        rc = lib().orc_rig_solve_normal(self._h, _p(params), C.c_double(lam), _p(step), _p(S), _p(gs))
        return rc, step, S[:ns, :ns], gs[:ns]

    def solve(self, params, mode=0, crit_type=3, max_count=200, eps=1e-7, policy=0, lambda0=1e-3, lambda_up=10.0,
              lambda_down=1.0 / 3.0, trace_cap=0):
        p = _d(params).copy()
        report = np.zeros(5)
        trace = np.zeros((max(trace_cap, 1), 5))
        rc = lib().orc_rig_solve(self._h, _p(p), int(mode), int(crit_type), int(max_count), C.c_double(eps),
                                 int(policy), C.c_double(lambda0), C.c_double(lambda_up), C.c_double(lambda_down),
                                 _p(report), _p(trace) if trace_cap else None, int(trace_cap))
        it = int(report[0])
        return dict(status=rc, params=p, iters=it, change=report[1], cost=report[2], lam=report[3],
                    trace=trace[:min(it, trace_cap)])

    def error(self, params, policy=0):
        out = np.zeros(5)
        per_edge = np.zeros(self.n_edge)
        lib().orc_rig_error(self._h, _p(_d(params)), int(policy), _p(out), _p(per_edge))
        return dict(mean_reproj_error=out[0], rms=out[1], sum_norm=out[2], sum_sq=out[3], n_points=int(out[4]),
                    per_edge=per_edge)


def omni_flags2idx(flags, n):
    idx = np.zeros(6 * n + 10, dtype=np.int32)
    lib().orc_omni_flags2idx(int(flags), int(n), _p(idx, c_int_p))
    return idx


def omni_solve(off, obj, img, param, flags=0, crit_type=3, max_count=200, eps=1e-8, dense=False, trace_cap=0):
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = off.size - 1
    obj = _d(obj).reshape(-1, 3); img = _d(img).reshape(-1, 2)
    p = _d(param).copy()
    assert p.size == 6 * n + 10
    report = np.zeros(4)
    trace = np.zeros((max(trace_cap, 1), 4))
    rc = lib().orc_omni_solve(n, _p(off, c_i64_p), _p(obj), _p(img), _p(p), int(flags), int(crit_type),
                              int(max_count), C.c_double(eps), int(dense), _p(report),
                              _p(trace) if trace_cap else None, int(trace_cap))
    it = int(report[0])
    return dict(status=rc, params=p, iters=it, change=report[1], cost=report[2], rms=report[3],
                trace=trace[:min(it, trace_cap)])


def omni_build(off, obj, img, param):
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = off.size - 1
    obj = _d(obj).reshape(-1, 3); img = _d(img).reshape(-1, 2)
    Hii = np.zeros((n, 6, 6)); HiI = np.zeros((n, 6, 10)); HII = np.zeros((10, 10)); gi = np.zeros((n, 6)); gI = np.zeros(10)
    cost = lib().orc_omni_build(n, _p(off, c_i64_p), _p(obj), _p(img), _p(_d(param)), _p(Hii), _p(HiI), _p(HII),
                                _p(gi), _p(gI))
    return dict(cost=float(cost), Hii=Hii, HiI=HiI, HII=HII, gi=gi, gI=gI)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))
