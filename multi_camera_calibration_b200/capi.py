"""ctypes binding of include/mccba.h (the C ABI of the CUDA core).  Plumbing only: every numerical operation happens
inside libmccba.so on the GPU.  There is no CPU fallback -- if the library or a CUDA device is missing this raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCCBA_LIB") or os.path.join(_HERE, "libmccba.so")   # MCCBA_LIB: development override (kernel variants)

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NUMERIC, ERR_NCCL = 0, 1, 2, 3, 4, 5
PINHOLE, OMNIDIRECTIONAL = 0, 1
CRIT_COUNT, CRIT_EPS = 1, 2
MODE_REFERENCE_GN, MODE_LM = 0, 1
PRECISION_FP64, PRECISION_MIXED, PRECISION_FAST32, PRECISION_AUTO = 0, 1, 2, 3


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("nccl_id", C.c_ubyte * 128),
                ("use_graph", C.c_int), ("verbose", C.c_int)]


class SolveOpts(C.Structure):
    _fields_ = [("mode", C.c_int), ("crit_type", C.c_int), ("max_count", C.c_int), ("epsilon", C.c_double),
                ("lambda0", C.c_double), ("lambda_up", C.c_double), ("lambda_down", C.c_double)]


class Report(C.Structure):
    _fields_ = [("iterations", C.c_int), ("accepted", C.c_int), ("rejected", C.c_int), ("status", C.c_int),
                ("graph_launches", C.c_int), ("kernel_launches", C.c_int), ("change", C.c_double), ("cost", C.c_double),
                ("lambda_", C.c_double), ("device_ms", C.c_double)]


class ErrorStats(C.Structure):
    _fields_ = [("mean_reproj_error", C.c_double), ("rms", C.c_double), ("sum_norm", C.c_double), ("sum_sq", C.c_double),
                ("n_points", C.c_int64)]


EXPORTS = ["mccba_exchange_mode", "mccba_default_options", "mccba_default_solve_opts", "mccba_nccl_unique_id", "mccba_create", "mccba_destroy",
           "mccba_last_error", "mccba_set_cameras", "mccba_set_observations", "mccba_set_parameters",
           "mccba_get_parameters", "mccba_save_parameters", "mccba_restore_parameters", "mccba_eval", "mccba_reduced_system", "mccba_solve", "mccba_reproj_error",
           "mccba_allreduce_sum", "mccba_last_kernel_ms", "mccba_time_eval", "mccba_debug_solve_dense", "mccba_omni_set_observations",
           "mccba_omni_set_parameters", "mccba_omni_get_parameters", "mccba_omni_solve", "mccba_omni_gram", "mccba_set_precision",
           "mccba_get_precision", "mccba_effective_precision", "mccba_stereo_set_observations", "mccba_stereo_set_parameters", "mccba_stereo_get_parameters",
           "mccba_stereo_solve", "mccba_stereo_uncertainties"]

_lib = None


class MccbaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mccba status %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libmccba.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m multi_camera_calibration_b200.build` "
                              "(nvcc, sm_100a); there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.mccba_last_error.restype = C.c_char_p
        _lib.mccba_last_error.argtypes = [C.c_void_p]
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def nccl_unique_id():
    buf = (C.c_ubyte * 128)()
    rc = lib().mccba_nccl_unique_id(buf)
    if rc:
        raise MccbaError(rc, "libnccl.so.2 not available")
    return bytes(buf)


class Solver:
    """One handle = one GPU + one stream (include/mccba.h)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_id=None, use_graph=True, precision=None):
        L = lib()
        o = Options()
        L.mccba_default_options(C.byref(o))
        o.device, o.rank, o.nranks, o.use_graph = int(device), int(rank), int(nranks), int(bool(use_graph))
        if nranks > 1:
            assert nccl_id is not None and len(nccl_id) == 128
            C.memmove(o.nccl_id, nccl_id, 128)
        self._h = C.c_void_p()
        rc = L.mccba_create(C.byref(o), C.byref(self._h))
        if rc:
            msg = L.mccba_last_error(self._h).decode() if self._h else "mccba_create failed (no CUDA device?)"
            if self._h:
                L.mccba_destroy(self._h)
                self._h = C.c_void_p()
            raise MccbaError(rc, msg)
        self.n_cam = self.n_frame = self.n_edge = 0
        self.nranks = nranks
        if precision is not None:
            self.set_precision(precision)

    def close(self):
        if getattr(self, "_h", None):
            lib().mccba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise MccbaError(rc, lib().mccba_last_error(self._h).decode())

    @property
    def n_param(self):
        return 6 * (self.n_cam + self.n_frame - 1)

    def set_cameras(self, model, K5, dist8, ndist, xi):
        model = np.ascontiguousarray(model, dtype=np.int32)
        n = model.size
        K5 = np.ascontiguousarray(K5, dtype=np.float64).reshape(n, 5)
        dist8 = np.ascontiguousarray(dist8, dtype=np.float64).reshape(n, 8)
        ndist = np.ascontiguousarray(ndist, dtype=np.int32).reshape(n)
        xi = np.ascontiguousarray(xi, dtype=np.float64).reshape(n)
        self._check(lib().mccba_set_cameras(self._h, n, _ptr(model, C.c_int), _ptr(K5, C.c_double),
                                            _ptr(dist8, C.c_double), _ptr(ndist, C.c_int), _ptr(xi, C.c_double)))
        self.n_cam = n

    def set_observations(self, n_frame, edge_cam, edge_pv, edge_off, obj_xyz, img_uv):
        edge_cam = np.ascontiguousarray(edge_cam, dtype=np.int32)
        edge_pv = np.ascontiguousarray(edge_pv, dtype=np.int32)
        edge_off = np.ascontiguousarray(edge_off, dtype=np.int64)
        obj = np.ascontiguousarray(obj_xyz, dtype=np.float32)
        img = np.ascontiguousarray(img_uv, dtype=np.float32)
        ne = edge_cam.size
        if edge_pv.size != ne or edge_off.size != ne + 1 or obj.size != 3 * int(edge_off[-1]) or img.size != 2 * int(edge_off[-1]):
            raise MccbaError(ERR_ARG, "set_observations: inconsistent array sizes")
        self._check(lib().mccba_set_observations(self._h, int(n_frame), ne, _ptr(edge_cam, C.c_int),
                                                 _ptr(edge_pv, C.c_int), _ptr(edge_off, C.c_int64),
                                                 _ptr(obj, C.c_float), _ptr(img, C.c_float)))
        self.n_frame, self.n_edge = int(n_frame), ne

    def set_rig(self, rig):
        """rig: dict in the C-ABI layout (tests/rigs.py, synth.make_rig)."""
        self.set_cameras(rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"])
        self.set_observations(rig["n_frame"], rig["edge_cam"], rig["edge_pv"], rig["edge_off"], rig["obj"], rig["img"])

    def set_precision(self, policy):
        """PRECISION_AUTO (3, default), PRECISION_FP64 (0), PRECISION_MIXED (1) or PRECISION_FAST32 (2); discards the current
        problem when the policy changes."""
        self._check(lib().mccba_set_precision(self._h, int(policy)))

    def get_precision(self):
        return int(lib().mccba_get_precision(self._h))

    def effective_precision(self):
        """What the last solve / evaluation ran: differs from get_precision() only under PRECISION_AUTO."""
        return int(lib().mccba_effective_precision(self._h))

    def set_parameters(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._check(lib().mccba_set_parameters(self._h, C.c_int64(p.size), _ptr(p, C.c_double)))

    def get_parameters(self, out=None):
        """Current parameters [rvec|tvec] per vertex 1..; `out` (float64, C-contiguous, n_param long, e.g. pinned memory)
        receives them without an intermediate allocation."""
        p = np.empty(self.n_param) if out is None else out
        assert p.dtype == np.float64 and p.flags["C_CONTIGUOUS"] and p.size == self.n_param
        self._check(lib().mccba_get_parameters(self._h, C.c_int64(p.size), _ptr(p, C.c_double)))
        return p

    def save_parameters(self):
        self._check(lib().mccba_save_parameters(self._h))

    def restore_parameters(self):
        self._check(lib().mccba_restore_parameters(self._h))

    def eval(self, want_blocks=True):
        cost = C.c_double()
        H6 = np.zeros((self.n_edge, 21)) if want_blocks else None
        g6 = np.zeros((self.n_edge, 6)) if want_blocks else None
        ec = np.zeros(self.n_edge) if want_blocks else None
        self._check(lib().mccba_eval(self._h, C.byref(cost), _ptr(H6, C.c_double), _ptr(g6, C.c_double),
                                     _ptr(ec, C.c_double)))
        return dict(cost=cost.value, H6=H6, g6=g6, edge_cost=ec)

    def reduced_system(self, lam=0.0):
        ns = 6 * (self.n_cam - 1)
        S = np.zeros((ns, ns)); gs = np.zeros(ns)
        self._check(lib().mccba_reduced_system(self._h, C.c_double(lam), _ptr(S, C.c_double), _ptr(gs, C.c_double)))
        return S, gs

    def solve(self, mode=MODE_REFERENCE_GN, crit_type=CRIT_COUNT, max_count=20, eps=1e-7, lambda0=1e-3, lambda_up=10.0,
              lambda_down=1.0 / 3.0, check=True):
        o = SolveOpts(int(mode), int(crit_type), int(max_count), float(eps), float(lambda0), float(lambda_up),
                      float(lambda_down))
        r = Report()
        rc = lib().mccba_solve(self._h, C.byref(o), C.byref(r))
        if rc and check:
            self._check(rc)
        return dict(rc=rc, iterations=r.iterations, accepted=r.accepted, rejected=r.rejected, status=r.status,
                    graph_launches=r.graph_launches, kernel_launches=r.kernel_launches, change=r.change, cost=r.cost,
                    lam=r.lambda_, device_ms=r.device_ms)

    def reproj_error(self):
        st = ErrorStats()
        pe = np.zeros(self.n_edge)
        self._check(lib().mccba_reproj_error(self._h, C.byref(st), _ptr(pe, C.c_double)))
        return dict(mean_reproj_error=st.mean_reproj_error, rms=st.rms, sum_norm=st.sum_norm, sum_sq=st.sum_sq,
                    n_points=int(st.n_points), per_edge=pe)

    def allreduce_sum(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).copy()
        self._check(lib().mccba_allreduce_sum(self._h, _ptr(a, C.c_double), int(a.size)))
        return a

    def last_kernel_ms(self):
        out = np.zeros(6)
        self._check(lib().mccba_last_kernel_ms(self._h, _ptr(out, C.c_double)))
        return out

    def exchange_mode(self):
        """0 single rank, 1 ncclAllReduce, 2 NVLink peer-memory windows (include/mccba.h: mccba_exchange_mode)."""
        return int(lib().mccba_exchange_mode(self._h))

    # ---- double-sided board calibration (include/mccba.h: mccba_ds_*) ----
    def ds_set_problem(self, edge_back, cam_pose):
        b = np.ascontiguousarray(edge_back, dtype=np.uint8)
        cp = np.ascontiguousarray(cam_pose, dtype=np.float64)
        self._check(lib().mccba_ds_set_problem(self._h, _ptr(b, C.c_ubyte), _ptr(cp, C.c_double)))

    def ds_set_parameters(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ds_n = p.size
        self._check(lib().mccba_ds_set_parameters(self._h, C.c_int64(p.size), _ptr(p, C.c_double)))

    def ds_get_parameters(self):
        p = np.zeros(self._ds_n)
        self._check(lib().mccba_ds_get_parameters(self._h, C.c_int64(p.size), _ptr(p, C.c_double)))
        return p

    def ds_solve(self, crit_type=3, max_count=200, eps=1e-7, check=True):
        r = Report()
        rc = lib().mccba_ds_solve(self._h, int(crit_type), int(max_count), C.c_double(eps), C.byref(r))
        if rc and check:
            self._check(rc)
        return dict(rc=rc, iterations=r.iterations, status=r.status, change=r.change, cost=r.cost, device_ms=r.device_ms,
                    kernel_launches=r.kernel_launches)

    def ds_normal(self):
        S = np.zeros((6, 6)); g = np.zeros(6); c = C.c_double(0.0)
        self._check(lib().mccba_ds_normal(self._h, _ptr(S, C.c_double), _ptr(g, C.c_double), C.byref(c)))
        return S, g, float(c.value)

    def exchange_stats(self):
        """Mean store / wait microseconds of the peer-memory exchange since the last call (mccba_exchange_stats)."""
        out = (C.c_double * 4)()
        self._check(lib().mccba_exchange_stats(self._h, out))
        return dict(store_us=out[0], wait_us=out[1], launches=int(out[2]), max_wait_us=out[3])

    def debug_solve_dense(self, S, g, blocked=True):
        S = np.ascontiguousarray(S, dtype=np.float64); g = np.ascontiguousarray(g, dtype=np.float64)
        n = g.size
        x = np.zeros(n)
        self._check(lib().mccba_debug_solve_dense(self._h, n, _ptr(S, C.c_double), _ptr(g, C.c_double),
                                                  _ptr(x, C.c_double), int(blocked)))
        return x, self.last_kernel_ms()[0]

    # ---- single-camera Mei calibration loop (cv::omnidir::calibrate without its closed-form initialisation) --------
    def omni_set_observations(self, frame_off, obj_xyz, img_uv):
        off = np.ascontiguousarray(frame_off, dtype=np.int64)
        obj = np.ascontiguousarray(obj_xyz, dtype=np.float32); img = np.ascontiguousarray(img_uv, dtype=np.float32)
        self._omni_n = off.size - 1
        self._omni_pts = int(off[-1])
        self._check(lib().mccba_omni_set_observations(self._h, int(self._omni_n), _ptr(off, C.c_int64), _ptr(obj, C.c_float),
                                                      _ptr(img, C.c_float)))
        if self.nranks > 1:      # frames shard over the ranks: the cost of a report is the whole job's, so is the corner count
            self._omni_pts = int(round(self.allreduce_sum([float(self._omni_pts)])[0]))

    def omni_set_parameters(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._check(lib().mccba_omni_set_parameters(self._h, C.c_int64(p.size), _ptr(p, C.c_double)))

    def omni_get_parameters(self):
        p = np.zeros(6 * self._omni_n + 10)
        self._check(lib().mccba_omni_get_parameters(self._h, C.c_int64(p.size), _ptr(p, C.c_double)))
        return p

    def omni_solve(self, flags=0, crit_type=3, max_count=200, eps=1e-8, check=True):
        r = Report()
        rc = lib().mccba_omni_solve(self._h, int(flags), int(crit_type), int(max_count), C.c_double(eps), C.byref(r))
        if rc and check:
            self._check(rc)
        return dict(rc=rc, iterations=r.iterations, status=r.status, change=r.change, cost=r.cost,
                    rms=float(np.sqrt(r.cost / self._omni_pts)), device_ms=r.device_ms, kernel_launches=r.kernel_launches)

    def omni_gram(self):
        g = np.zeros((self._omni_n, 17, 17)); cost = C.c_double()
        self._check(lib().mccba_omni_gram(self._h, _ptr(g, C.c_double), C.byref(cost)))
        return g, cost.value

    # ---- omnidir stereo bundle adjustment (cv::omnidir::stereoCalibrate's loop + estimateUncertaintiesStereo) ----------
    def stereo_set_observations(self, frame_off, obj_xyz, img1_uv, img2_uv):
        off = np.ascontiguousarray(frame_off, dtype=np.int64)
        obj = np.ascontiguousarray(obj_xyz, dtype=np.float32)
        i1 = np.ascontiguousarray(img1_uv, dtype=np.float32); i2 = np.ascontiguousarray(img2_uv, dtype=np.float32)
        self._st_n = off.size - 1
        self._st_pts = int(off[-1])
        self._check(lib().mccba_stereo_set_observations(self._h, int(self._st_n), _ptr(off, C.c_int64), _ptr(obj, C.c_float),
                                                        _ptr(i1, C.c_float), _ptr(i2, C.c_float)))

    def stereo_set_parameters(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._check(lib().mccba_stereo_set_parameters(self._h, C.c_int64(p.size), _ptr(p, C.c_double)))

    def stereo_get_parameters(self):
        p = np.zeros(6 * (self._st_n + 1) + 20)
        self._check(lib().mccba_stereo_get_parameters(self._h, C.c_int64(p.size), _ptr(p, C.c_double)))
        return p

    def stereo_solve(self, flags=0, crit_type=3, max_count=200, eps=1e-6, check=True):
        r = Report()
        rc = lib().mccba_stereo_solve(self._h, int(flags), int(crit_type), int(max_count), C.c_double(eps), C.byref(r))
        if rc and check:
            self._check(rc)
        return dict(rc=rc, iterations=r.iterations, status=r.status, change=r.change, cost=r.cost,
                    rms=float(np.sqrt(r.cost / (2 * self._st_pts))), device_ms=r.device_ms, kernel_launches=r.kernel_launches)

    def stereo_uncertainties(self, flags=0):
        e = np.zeros(6 * (self._st_n + 1) + 20); sd = np.zeros(2); rms = C.c_double()
        self._check(lib().mccba_stereo_uncertainties(self._h, int(flags), _ptr(e, C.c_double), _ptr(sd, C.c_double), C.byref(rms)))
        return dict(errors=e, std_error=sd, rms=rms.value)

    def time_eval(self, reps=10):
        ms = C.c_double()
        self._check(lib().mccba_time_eval(self._h, int(reps), C.byref(ms)))
        return ms.value
