"""Double-sided board calibration (src/doubleSide.cpp: cameras fixed, front<->back transform D + one
pose per frame).  CPU: the oracle restatement (oracle/dense_reenact.py: ds_*) against finite differences and against the
ground truth of a synthetic problem.  GPU: mccba_ds_* through the C ABI against the oracle -- iterates of the reference
schedule, the converged parameters and the cost -- under both the FP64 and the default precision policy."""
import numpy as np
import pytest

from multi_camera_calibration_b200 import synth
from oracle import dense_reenact as dr


def _problem(n_frame, seed=4001, **kw):
    r = synth.make_double_side_rig(n_frame, seed=seed, **kw)
    nC = r["n_cam"]
    K = np.zeros((nC, 3, 3))
    for c in range(nC):
        fx, fy, cx, cy, sk = r["cam_K5"][c]
        K[c] = [[fx, sk, cx], [0, fy, cy], [0, 0, 1]]
    dist = [r["cam_dist8"][c][:r["cam_ndist"][c]] for c in range(nC)]
    off = r["edge_off"]
    edges = [(int(r["edge_cam"][e]), int(r["edge_pv"][e]), r["obj"][off[e]:off[e + 1]], r["img"][off[e]:off[e + 1]])
             for e in range(r["edge_cam"].size)]
    return r, dr.RigProblem([0] * nC, K, dist, np.zeros(nC), edges, nC + r["n_frame"])


def test_oracle_jacobian_matches_finite_differences():
    r, prob = _problem(6)
    p0 = r["ds_params_init"]
    _, J, E, _, _ = dr.ds_compute_jacobian(prob, r["edge_back"], r["cam_pose"], p0, dense_out=True)
    h = 1e-6
    for k in range(p0.size):
        pp, pm = p0.copy(), p0.copy()
        pp[k] += h; pm[k] -= h
        _, _, Ep, _, _ = dr.ds_compute_jacobian(prob, r["edge_back"], r["cam_pose"], pp, dense_out=True)
        _, _, Em, _, _ = dr.ds_compute_jacobian(prob, r["edge_back"], r["cam_pose"], pm, dense_out=True)
        fd = -(Ep - Em) / (2 * h)          # E = observed - projected, J = d projected / d parameter
        assert np.abs(fd - J[:, k]).max() <= 2e-6 * max(np.abs(J[:, k]).max(), 1.0), k
    # front edges do not depend on D (src/doubleSide.cpp:334-335)
    loc = np.concatenate([[0], np.cumsum([2 * e[2].shape[0] for e in prob.edges])])
    for e in range(len(prob.edges)):
        if not r["edge_back"][e]:
            assert not J[loc[e]:loc[e + 1], 0:6].any()


def test_oracle_recovers_the_transform():
    r, prob = _problem(12)
    p, it, change = dr.ds_optimize(prob, r["edge_back"], r["cam_pose"], r["ds_params_init"], 3, 60, 1e-9)
    assert it < 60 and change <= 1e-9
    rms = np.sqrt(dr.ds_cost(prob, r["edge_back"], r["cam_pose"], p) / r["n_points"])
    assert 0.25 < rms < 0.45                                   # 0.3 px noise per coordinate
    d = np.abs(p[:6] - r["ds_params_true"][:6])
    assert d[:3].max() < 5e-3 and d[3:].max() < 0.5            # rad, mm


@pytest.fixture(scope="module")
def gpu_case():
    import multi_camera_calibration_b200 as m
    r, prob = _problem(40)
    solvers = {}
    for name, prec in (("fp64", m.capi.PRECISION_FP64), ("mixed", m.capi.PRECISION_MIXED)):
        s = m.Solver(device=0, precision=prec)
        s.set_rig(r)
        s.ds_set_problem(r["edge_back"], r["cam_pose"])
        solvers[name] = s
    yield r, prob, solvers
    for s in solvers.values():
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("policy,tol", [("fp64", 1e-9), ("mixed", 1e-6)])
def test_gpu_iterates_match_oracle(gpu_case, policy, tol):
    r, prob, solvers = gpu_case
    s = solvers[policy]
    rec = []
    dr.ds_optimize(prob, r["edge_back"], r["cam_pose"], r["ds_params_init"], 1, 4, 0.0, record=rec)
    for k in (1, 2, 4):
        s.ds_set_parameters(r["ds_params_init"])
        rep = s.ds_solve(1, k, 0.0)
        p = s.ds_get_parameters()
        ref = rec[k - 1]["params"]
        assert rep["iterations"] == k
        assert np.max(np.abs(p - ref) / np.maximum(np.abs(ref), 1.0)) < tol, (policy, k)
        assert abs(rep["change"] - rec[k - 1]["change"]) <= max(tol, 1e-9) * max(rec[k - 1]["change"], 1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("policy,tol", [("fp64", 1e-8), ("mixed", 1e-6)])
def test_gpu_converged_matches_oracle(gpu_case, policy, tol):
    r, prob, solvers = gpu_case
    s = solvers[policy]
    s.ds_set_parameters(r["ds_params_init"])
    rep = s.ds_solve(3, 200, 1e-7)                            # TermCriteria(COUNT + EPS, 200, 1e-7), mymulticalib.hpp:95
    ref, it, change = dr.ds_optimize(prob, r["edge_back"], r["cam_pose"], r["ds_params_init"], 3, 200, 1e-7)
    p = s.ds_get_parameters()
    assert rep["iterations"] == it
    assert np.max(np.abs(p - ref) / np.maximum(np.abs(ref), 1.0)) < tol
    cost = dr.ds_cost(prob, r["edge_back"], r["cam_pose"], ref)
    assert abs(rep["cost"] - cost) <= 1e-6 * cost
    d = np.abs(p[:6] - r["ds_params_true"][:6])
    assert d[:3].max() < 5e-3 and d[3:].max() < 0.5


@pytest.mark.gpu
def test_gpu_reduced_system_is_the_schur_complement(gpu_case):
    """mccba_ds_normal: the 6 x 6 system of D after elimination of the frame poses.  The oracle's J^T J is in additive
    Rodrigues coordinates; the GPU works in left-perturbation tangents, psi = J_l(om) d(om), so S_t = A^-T S_o A^-1 with
    A = blockdiag(J_l(om_D), I)."""
    import cv2
    r, prob, solvers = gpu_case
    s = solvers["fp64"]
    p0 = r["ds_params_init"]
    s.ds_set_parameters(p0)
    S, g, cost = s.ds_normal()
    _, J, E, JTJ, JTE = dr.ds_compute_jacobian(prob, r["edge_back"], r["cam_pose"], p0, dense_out=True)
    A, Bm, Cm = JTJ[:6, :6], JTJ[:6, 6:], JTJ[6:, 6:]
    So = A - Bm @ np.linalg.solve(Cm, Bm.T)
    go = JTE[:6] - Bm @ np.linalg.solve(Cm, JTE[6:])
    # left Jacobian of SO(3) at om_D from finite differences of log(exp(om + d) exp(om)^T)
    om = p0[:3]
    R0 = cv2.Rodrigues(om.reshape(3, 1))[0]
    Jl = np.zeros((3, 3))
    h = 1e-6
    for k in range(3):
        d = np.zeros(3); d[k] = h
        Rp = cv2.Rodrigues((om + d).reshape(3, 1))[0]; Rm = cv2.Rodrigues((om - d).reshape(3, 1))[0]
        # log of a rotation by ~1e-6 rad from its skew part (cv2.Rodrigues returns 0 below sin(theta) = 1e-5)
        vee = lambda M: 0.5 * np.array([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
        Jl[:, k] = (vee(Rp @ R0.T) - vee(Rm @ R0.T)) / (2 * h)
    Ainv = np.linalg.inv(np.block([[Jl, np.zeros((3, 3))], [np.zeros((3, 3)), np.eye(3)]]))
    St = Ainv.T @ So @ Ainv
    gt = Ainv.T @ go
    assert abs(cost - float(E @ E)) <= 1e-10 * cost
    assert np.abs(S - St).max() <= 1e-6 * np.abs(St).max()          # finite-difference J_l: 1e-6 is its accuracy
    assert np.abs(g - gt).max() <= 1e-6 * np.abs(gt).max()


@pytest.mark.gpu
def test_gpu_rejects_unobservable_transform(gpu_case):
    import multi_camera_calibration_b200 as m
    r, prob, solvers = gpu_case
    s = m.Solver(device=0)
    s.set_rig(r)
    with pytest.raises(m.MccbaError):
        s.ds_set_problem(np.zeros_like(r["edge_back"]), r["cam_pose"])      # no back edge: D is not observable
    with pytest.raises(m.MccbaError):
        s.ds_solve(1, 1, 0.0)                                                # no problem set
    s.close()


# ---- the host class: reference directory layout -> DoubleSideCalibration -> doublesideTransform.yaml ---------------------
SERIALS = ["839112060578", "839512061262", "f0220380"]


def _write_ds_dataset(tmp, r):
    import os
    cv2 = pytest.importorskip("cv2")
    data = os.path.join(tmp, "color"); cfg = os.path.join(tmp, "configs")
    os.makedirs(cfg)
    for c in range(r["n_cam"]):
        os.makedirs(os.path.join(data, SERIALS[c]))
        fs = cv2.FileStorage(os.path.join(cfg, SERIALS[c] + ".xml"), cv2.FILE_STORAGE_WRITE)
        fx, fy, cx, cy, _ = r["cam_K5"][c]
        pose = np.eye(4); pose[:3, :3] = r["cam_R"][c]; pose[:3, 3] = r["cam_t"][c]
        fs.write("CameraMatrix", pose)                     # what writeParameters2config wrote (src/mymulticalib.cpp:425-454)
        fs.write("Intrinsics", np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64))
        fs.write("Distortion", r["cam_dist8"][c][:5].reshape(1, 5).astype(np.float64))
        fs.release()
    off = r["edge_off"]
    for e in range(r["edge_cam"].size):
        c = int(r["edge_cam"][e]); ts = int(r["edge_pv"][e]) - r["n_cam"] + 100      # three digits: cv::glob (string) order = numeric order
        fs = cv2.FileStorage(os.path.join(data, SERIALS[c], "%d.yaml" % ts), cv2.FILE_STORAGE_WRITE)
        fs.write("corners", r["img"][off[e]:off[e + 1]].astype(np.float64))
        fs.write("objects", r["obj"][off[e]:off[e + 1]].astype(np.float64))
        fs.release()
    return data, cfg


def _host_case(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from multi_camera_calibration_b200 import multicalib
    r, prob = _problem(24, back_shape=(8, 5, 40.0))
    data, cfg = _write_ds_dataset(str(tmp_path), r)
    calib = multicalib.DoubleSideCalibration(SERIALS, multicalib.PINHOLE, 3, data, cfg, (9, 6), (8, 5), criteria=(3, 200, 1e-8))
    calib.loadImages()
    calib.initialize()
    return cv2, r, prob, calib


def test_host_class_loads_both_sides_and_initialises(tmp_path):
    cv2, r, prob, calib = _host_case(tmp_path)
    idx = calib.indexing()
    # both pattern sides are kept (src/doubleSide.cpp:114-118); MyMultiCameraCalibration would drop the 8 x 5 images
    assert idx["vertex_timestamp"].size == 3 + r["n_frame"] and idx["edge_cam"].size == r["edge_cam"].size
    assert np.array_equal(idx["edge_cam"], r["edge_cam"]) and np.array_equal(idx["edge_pv"], r["edge_pv"])
    T0 = calib.doubleSideTransform()
    Rd = cv2.Rodrigues(r["ds_params_true"][:3].reshape(3, 1))[0]
    assert np.abs(T0[:3, :3] - Rd).max() < 2e-2 and np.abs(T0[:3, 3] - r["ds_params_true"][3:6]).max() < 15.0   # PnP accuracy
    calib.close()


@pytest.mark.gpu
def test_host_class_from_the_reference_directory_layout(tmp_path):
    import os
    cv2, r, prob, calib = _host_case(tmp_path)
    rms = calib.optimizeExtrinsics()
    T = calib.doubleSideTransform()
    # the same minimum as the oracle reaches from its own (different) starting point
    ref, it, _ = dr.ds_optimize(prob, r["edge_back"], r["cam_pose"], r["ds_params_init"], 3, 200, 1e-9)
    Rref = cv2.Rodrigues(ref[:3].reshape(3, 1))[0]
    assert np.abs(T[:3, :3] - Rref).max() < 2e-6 and np.abs(T[:3, 3] - ref[3:6]).max() < 2e-3
    assert abs(rms - np.sqrt(dr.ds_cost(prob, r["edge_back"], r["cam_pose"], ref) / r["n_points"])) < 1e-5
    out = os.path.join(str(tmp_path), "doublesideTransform.yaml")
    calib.writeParameters(out)
    fs = cv2.FileStorage(out, cv2.FILE_STORAGE_READ)                    # MyMultiCameraCalibration::readDoubleSide (:98-104)
    Tr = fs.getNode("transform").mat()
    fs.release()
    assert Tr.shape == (4, 4) and np.abs(Tr - T).max() < 1e-12
    calib.close()
