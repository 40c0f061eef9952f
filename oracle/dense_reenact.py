"""TEST INFRASTRUCTURE ONLY -- tier-2 oracle: dense, literal numpy re-enactment of the reference loops.

Nothing under oracle/ is part of the product.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package.

This module re-enacts, line by line and at toy sizes only, the reference's *dense* formulation:
it builds the full 2M x P Jacobian, forms J^T J / J^T E densely and solves the full P x P system,
calling cv2 exactly where the reference calls OpenCV.  It exists to pin that the block-sparse /
Schur C oracle (oracle/mccba_oracle.c) and the CUDA path compute the same iterates.

PARITY STATUS: parity unpinned by the reference's own tests (it has none) and the
reference cannot be compiled here (no OpenCV C++ headers, no Eigen).  Pins used instead:
  tier 0  cv2.projectPoints / cv2.Rodrigues / cv2.composeRT / cv2.matMulDeriv (third-party truth)
  tier 1  the transcription of src/omnidir.cpp:126-243 below, finite-difference checked in tests
  tier 2  this dense re-enactment

Reference citations (relative to /root/reference):
  optimize_extrinsics            src/multicalib.cpp:462-514
  compute_jacobian_extrinsic     src/multicalib.cpp:593-703
  compute_photo_camera_jacobian  src/multicalib.cpp:717-824
  compose_motion                 src/multicalib.cpp:1008-1056
  compute_project_error          src/multicalib.cpp:895-1006
  omnidir_project_points         src/omnidir.cpp:84-245
  omni_compute_jacobian          src/omnidir.cpp:851-935
  omni_calibrate_loop            src/omnidir.cpp:1119-1147
  flags2idx                      src/omnidir.cpp:2031-2076
  omni_rms                       src/omnidir.cpp:1794-1802
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is third-party truth for tier 0; present in this image (4.13.0)
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

PINHOLE = 0
OMNIDIRECTIONAL = 1

CALIB_USE_GUESS = 1
CALIB_FIX_SKEW = 2
CALIB_FIX_K1 = 4
CALIB_FIX_K2 = 8
CALIB_FIX_P1 = 16
CALIB_FIX_P2 = 32
CALIB_FIX_XI = 64
CALIB_FIX_GAMMA = 128
CALIB_FIX_CENTER = 256


# --------------------------------------------------------------------------------------------
# Mei projection + 2x16 Jacobian: transcription of src/omnidir.cpp:126-243
# --------------------------------------------------------------------------------------------
def omnidir_project_points(obj, rvec, tvec, K, xi, D, want_jac=True):
    """obj (N,3) float64; returns proj (N,2) float64 and jac (2N,16) float64.
    Column layout (src/omnidir.cpp:65-73): dom(3) dT(3) df(2) ds(1) dc(2) dxi(1) dkp(4)."""
    obj = np.asarray(obj, dtype=np.float64).reshape(-1, 3)
    om = np.asarray(rvec, dtype=np.float64).reshape(3)
    T = np.asarray(tvec, dtype=np.float64).reshape(3)
    K = np.asarray(K, dtype=np.float64)
    f = np.array([K[0, 0], K[1, 1]])
    c = np.array([K[0, 2], K[1, 2]])
    s = K[0, 1]
    k1, k2, p1, p2 = [float(v) for v in np.asarray(D, dtype=np.float64).reshape(-1)[:4]]
    R, dRdom = cv2.Rodrigues(om)  # dRdom is 3x9 (rows = d/dom_k)
    n = obj.shape[0]
    proj = np.zeros((n, 2))
    jac = np.zeros((2 * n, 16)) if want_jac else None
    for i in range(n):
        Xw = obj[i]
        Xc = R @ Xw + T
        nrm = np.linalg.norm(Xc)
        Xs = Xc / nrm
        xu = np.array([Xs[0] / (Xs[2] + xi), Xs[1] / (Xs[2] + xi)])
        r2 = xu[0] * xu[0] + xu[1] * xu[1]
        r4 = r2 * r2
        xd = np.array([
            xu[0] * (1 + k1 * r2 + k2 * r4) + 2 * p1 * xu[0] * xu[1] + p2 * (r2 + 2 * xu[0] * xu[0]),
            xu[1] * (1 + k1 * r2 + k2 * r4) + p1 * (r2 + 2 * xu[1] * xu[1]) + 2 * p2 * xu[0] * xu[1]])
        proj[i, 0] = f[0] * xd[0] + s * xd[1] + c[0]
        proj[i, 1] = f[1] * xd[1] + c[1]
        if not want_jac:
            continue
        dXcdR = np.zeros((3, 9))
        dXcdR[0, 0:3] = Xw
        dXcdR[1, 3:6] = Xw
        dXcdR[2, 6:9] = Xw
        dXcdom = dXcdR @ dRdom.T
        r_1 = 1.0 / nrm
        r_3 = r_1 ** 3
        dXsdXc = np.array([
            [r_1 - Xc[0] * Xc[0] * r_3, -(Xc[0] * Xc[1]) * r_3, -(Xc[0] * Xc[2]) * r_3],
            [-(Xc[0] * Xc[1]) * r_3, r_1 - Xc[1] * Xc[1] * r_3, -(Xc[1] * Xc[2]) * r_3],
            [-(Xc[0] * Xc[2]) * r_3, -(Xc[1] * Xc[2]) * r_3, r_1 - Xc[2] * Xc[2] * r_3]])
        den = Xs[2] + xi
        dxudXs = np.array([[1 / den, 0, -Xs[0] / den / den],
                           [0, 1 / den, -Xs[1] / den / den]])
        temp1 = 2 * k1 * xu[0] + 4 * k2 * xu[0] * r2
        temp2 = 2 * k1 * xu[1] + 4 * k2 * xu[1] * r2
        dxddxu = np.array([
            [k2 * r4 + 6 * p2 * xu[0] + 2 * p1 * xu[1] + xu[0] * temp1 + k1 * r2 + 1,
             2 * p1 * xu[0] + 2 * p2 * xu[1] + xu[0] * temp2],
            [2 * p1 * xu[0] + 2 * p2 * xu[1] + xu[1] * temp1,
             k2 * r4 + 2 * p2 * xu[0] + 6 * p1 * xu[1] + xu[1] * temp2 + k1 * r2 + 1]])
        dxpddxd = np.array([[f[0], s], [0, f[1]]])
        dxpddXc = dxpddxd @ dxddxu @ dxudXs @ dXsdXc
        dxpddom = dxpddXc @ dXcdom
        dxpddT = dxpddXc
        dxudxi = np.array([[-Xs[0] / den / den], [-Xs[1] / den / den]])
        dxpddxi = dxpddxd @ dxddxu @ dxudxi
        dxddkp = np.array([
            [xu[0] * r2, xu[0] * r4, 2 * xu[0] * xu[1], r2 + 2 * xu[0] * xu[0]],
            [xu[1] * r2, xu[1] * r4, r2 + 2 * xu[1] * xu[1], 2 * xu[0] * xu[1]]])
        dxpddkp = dxpddxd @ dxddkp
        row = jac[2 * i:2 * i + 2]
        row[:, 0:3] = dxpddom
        row[:, 3:6] = dxpddT
        row[0, 6] = xd[0]; row[1, 7] = xd[1]          # df
        row[0, 8] = xd[1]; row[1, 8] = 0.0            # ds
        row[0, 9] = 1.0; row[1, 10] = 1.0             # dc
        row[:, 11] = dxpddxi[:, 0]
        row[:, 12:16] = dxpddkp
    return proj, jac


# --------------------------------------------------------------------------------------------
# compose_motion: src/multicalib.cpp:1008-1056 (inputs widened to double at :1013-1016)
# --------------------------------------------------------------------------------------------
def compose_motion(om1, T1, om2, T2):
    om1 = np.asarray(om1, dtype=np.float64).reshape(3, 1)
    om2 = np.asarray(om2, dtype=np.float64).reshape(3, 1)
    T1 = np.asarray(T1, dtype=np.float64).reshape(3, 1)
    T2 = np.asarray(T2, dtype=np.float64).reshape(3, 1)
    R1, dR1dom1 = cv2.Rodrigues(om1)
    R2, dR2dom2 = cv2.Rodrigues(om2)
    dR1dom1 = dR1dom1.T  # 9x3
    dR2dom2 = dR2dom2.T
    R3 = R2 @ R1
    dR3dR2, dR3dR1 = cv2.matMulDeriv(R2, R1)
    om3, dom3dR3 = cv2.Rodrigues(R3)
    dom3dR3 = dom3dR3.T  # 3x9
    dom3dom1 = dom3dR3 @ dR3dR1 @ dR1dom1
    dom3dom2 = dom3dR3 @ dR3dR2 @ dR2dom2
    dom3dT1 = np.zeros((3, 3)); dom3dT2 = np.zeros((3, 3))
    T3t = R2 @ T1
    dT3tdR2, dT3tdT1 = cv2.matMulDeriv(R2, T1)
    dT3tdom2 = dT3tdR2 @ dR2dom2
    T3 = T3t + T2
    return (om3.reshape(3), T3.reshape(3), dom3dom1, dom3dT1, dom3dom2, dom3dT2,
            np.zeros((3, 3)), dT3tdT1, dT3tdom2, np.eye(3))


# --------------------------------------------------------------------------------------------
# Problem container (follows the reference's members, include/opencv2/ccalib/multicalib.hpp:193-214)
# --------------------------------------------------------------------------------------------
class RigProblem:
    """cam_type: PINHOLE or OMNIDIRECTIONAL per camera (the reference has one global type, fact 7).
    K: (nC,3,3), dist: list of 1-D arrays (pinhole: >=4, omnidir: exactly 4), xi: (nC,).
    edges: list of (cameraVertex, photoVertex, obj (N,3) f32, img (N,2) f32) in the reference's edge order.
    n_vertex = nC + number of photo vertices.  params: 6*(n_vertex-1) = [rvec|tvec] per vertex 1..nV-1."""

    def __init__(self, cam_type, K, dist, xi, edges, n_vertex):
        self.cam_type = list(cam_type)
        self.K = np.asarray(K, dtype=np.float64)
        self.dist = [np.asarray(d, dtype=np.float64).reshape(-1) for d in dist]
        self.xi = np.asarray(xi, dtype=np.float64).reshape(-1)
        self.edges = edges
        self.n_camera = len(self.cam_type)
        self.n_vertex = int(n_vertex)

    @property
    def n_param(self):
        return 6 * (self.n_vertex - 1)


def _project_with_jac(prob, cam, obj, rvec, tvec):
    if prob.cam_type[cam] == PINHOLE:
        Kp = prob.K[cam].copy()
        proj, jac = cv2.projectPoints(obj.reshape(-1, 1, 3), rvec.reshape(3, 1), tvec.reshape(3, 1), Kp,
                                      prob.dist[cam].reshape(1, -1))
        return proj.reshape(-1, 2), np.asarray(jac, dtype=np.float64)[:, 0:6]
    proj, jac = omnidir_project_points(obj, rvec, tvec, prob.K[cam], float(prob.xi[cam]), prob.dist[cam])
    return proj, jac[:, 0:6]


def compute_photo_camera_jacobian(prob, e, rvecP, tvecP, rvecC, tvecC, policy):
    """src/multicalib.cpp:717-824.  policy: 'fp64' or 'faithful_f32' (SURVEY.md appendix A)."""
    cam, _, obj32, img32 = prob.edges[e]
    (om3, T3, dom3dom1, dom3dT1, dom3dom2, dom3dT2,
     dT3dom1, dT3dT1, dT3dom2, dT3dT2) = compose_motion(rvecP, tvecP, rvecC, tvecC)
    if policy == 'faithful_f32':
        om3 = om3.astype(np.float32)              # :742-749
        T3 = T3.astype(np.float32)
        obj = obj32.astype(np.float32)
        proj, jac = _project_with_jac(prob, cam, obj, om3.astype(np.float64) if prob.cam_type[cam] else om3,
                                      T3.astype(np.float64) if prob.cam_type[cam] else T3)
        proj = proj.astype(np.float32)            # output depth follows objectPoints
        E = (img32.astype(np.float32) - proj).astype(np.float64)   # :789-792 float32 subtraction
    else:
        obj = obj32.astype(np.float64)
        proj, jac = _project_with_jac(prob, cam, obj, om3, T3)
        E = img32.astype(np.float64) - proj
    E = E.reshape(-1)                             # interleaved (x0,y0,x1,y1,...) :797
    # camera pose is (om2,T2), photo pose is (om1,T1) at the call site :734
    dx_drC = jac[:, 0:3] @ dom3dom2 + jac[:, 3:6] @ dT3dom2
    dx_dtC = jac[:, 0:3] @ dom3dT2 + jac[:, 3:6] @ dT3dT2
    dx_drP = jac[:, 0:3] @ dom3dom1 + jac[:, 3:6] @ dT3dom1
    dx_dtP = jac[:, 0:3] @ dom3dT1 + jac[:, 3:6] @ dT3dT1
    return np.hstack([dx_drP, dx_dtP]), np.hstack([dx_drC, dx_dtC]), E


def compute_jacobian_extrinsic(prob, params, policy, dense_out=False):
    """src/multicalib.cpp:593-703 -- dense J (2M x P), JTJ, JTE, and the full-system solve."""
    nE = len(prob.edges)
    loc = np.zeros(nE + 1, dtype=np.int64)
    for e in range(nE):
        loc[e + 1] = loc[e] + 2 * prob.edges[e][2].shape[0]          # :597-603
    P = prob.n_param
    J = np.zeros((loc[nE], P))
    E = np.zeros(loc[nE])
    for e in range(nE):
        cam, pv = prob.edges[e][0], prob.edges[e][1]
        rvecP = params[(pv - 1) * 6:(pv - 1) * 6 + 3]
        tvecP = params[(pv - 1) * 6 + 3:(pv - 1) * 6 + 6]
        if cam > 0:
            rvecC = params[(cam - 1) * 6:(cam - 1) * 6 + 3]
            tvecC = params[(cam - 1) * 6 + 3:(cam - 1) * 6 + 6]
        else:
            rvecC = np.zeros(3, dtype=params.dtype)
            tvecC = np.zeros(3, dtype=params.dtype)
        Jp, Jc, err = compute_photo_camera_jacobian(prob, e, rvecP, tvecP, rvecC, tvecC, policy)
        if cam > 0:
            J[loc[e]:loc[e + 1], (cam - 1) * 6:cam * 6] = Jc
        J[loc[e]:loc[e + 1], (pv - 1) * 6:pv * 6] = Jp
        E[loc[e]:loc[e + 1]] = err
    JTJ = J.T @ J                                                       # :688
    JTE = J.T @ E                                                       # :689
    # Eigen CG on a SPD system converges to the exact solution; direct solve here.
    x = np.linalg.solve(JTJ, JTE)
    if dense_out:
        return x, J, E, JTJ, JTE
    return x, float(E @ E), loc


def optimize_extrinsics(prob, params0, crit_type, max_count, eps, policy='fp64', record=None):
    """src/multicalib.cpp:462-514.  Returns (params, n_iter, change)."""
    if policy == 'faithful_f32':
        params = np.asarray(params0, dtype=np.float32).copy()
    else:
        params = np.asarray(params0, dtype=np.float64).copy()
    change = 1.0
    it = 0
    while True:
        if ((crit_type == 1 and it >= max_count) or (crit_type == 2 and change <= eps) or
                (crit_type == 3 and (change <= eps or it >= max_count))):
            break
        alpha = 0.95 ** (it + 1.0)                                      # :482-483
        x, cost, _ = compute_jacobian_extrinsic(prob, params, policy)
        G = alpha * x
        if policy == 'faithful_f32':
            G = G.astype(np.float32)                                    # :493-497
            params = params + G                                         # float32 add :501
        else:
            params = params + G
        change = float(np.linalg.norm(G.astype(np.float64)) / np.linalg.norm(params.astype(np.float64)))
        if record is not None:
            record.append(dict(iter=it, cost_before=cost, change=change, params=params.astype(np.float64).copy()))
        it += 1
    return params.astype(np.float64), it, change


# --------------------------------------------------------------------------------------------
# Double-sided board calibration: src/doubleSide.cpp (cameras fixed, unknown = front<->back transform D + frame poses)
# --------------------------------------------------------------------------------------------
def ds_photo_doubleside_jacobian(prob, e, back, rvecP, tvecP, rvecC, tvecC, rvecD, tvecD):
    """src/doubleSide.cpp:288-429 (fp64 policy).  Returns (jacobianPhoto (2N,6), jacobianDoubleside (2N,6), E (2N,))."""
    cam, _, obj32, img32 = prob.edges[e]
    # front pose = camera o photo (:312-314): om1 = photo, om2 = camera
    (omF, TF, dRF_drP, dRF_dtP, _, _, dTF_drP, dTF_dtP, _, _) = compose_motion(rvecP, tvecP, rvecC, tvecC)
    if back:
        # back pose = front o doubleside (:318-321): om1 = doubleside, om2 = front
        (omT, TT, dRT_drD, dRT_dtD, dRT_drF, dRT_dtF, dTT_drD, dTT_dtD, dTT_drF, dTT_dtF) = compose_motion(rvecD, tvecD, omF, TF)
        dRT_drP = dRT_drF @ dRF_drP                  # :324-327
        dRT_dtP = dRT_dtF @ dTF_dtP
        dTT_drP = dTT_drF @ dRF_drP
        dTT_dtP = dTT_dtF @ dTF_dtP
    else:
        omT, TT = omF, TF                            # :328-338
        dRT_drP, dRT_dtP, dTT_drP, dTT_dtP = dRF_drP, dRF_dtP, dTF_drP, dTF_dtP
        dRT_drD = dRT_dtD = dTT_drD = dTT_dtD = np.zeros((3, 3))
    obj = obj32.astype(np.float64)
    proj, jac = _project_with_jac(prob, cam, obj, omT, TT)
    E = (img32.astype(np.float64) - proj).reshape(-1)
    dx_drD = jac[:, 0:3] @ dRT_drD + jac[:, 3:6] @ dTT_drD      # :393-394
    dx_dtD = jac[:, 0:3] @ dRT_dtD + jac[:, 3:6] @ dTT_dtD
    dx_drP = jac[:, 0:3] @ dRT_drP + jac[:, 3:6] @ dTT_drP      # :409-410
    dx_dtP = jac[:, 0:3] @ dRT_dtP + jac[:, 3:6] @ dTT_dtP
    return np.hstack([dx_drP, dx_dtP]), np.hstack([dx_drD, dx_dtD]), E


def ds_compute_jacobian(prob, back, cam_pose, params, dense_out=False):
    """src/doubleSide.cpp:434-581.  params = [D | photo vertices in vertex order] (buildParas :233-261); the column block
    of photo vertex pv is pv - nCamera + 1 (getPhotoVertexParameters :431-433).  cam_pose: (nC, 6) fixed [rvec | tvec]."""
    nE = len(prob.edges)
    nC = prob.n_camera
    loc = np.zeros(nE + 1, dtype=np.int64)
    for e in range(nE):
        loc[e + 1] = loc[e] + 2 * prob.edges[e][2].shape[0]
    P = params.size
    J = np.zeros((loc[nE], P))
    E = np.zeros(loc[nE])
    rvecD, tvecD = params[0:3], params[3:6]
    for e in range(nE):
        cam, pv = prob.edges[e][0], prob.edges[e][1]
        row = pv - nC + 1
        rvecP, tvecP = params[row * 6:row * 6 + 3], params[row * 6 + 3:row * 6 + 6]
        Jp, Jd, err = ds_photo_doubleside_jacobian(prob, e, bool(back[e]), rvecP, tvecP, cam_pose[cam, 0:3], cam_pose[cam, 3:6], rvecD, tvecD)
        J[loc[e]:loc[e + 1], 0:6] = Jd                                  # :544-545
        J[loc[e]:loc[e + 1], row * 6:(row + 1) * 6] = Jp                # :547-548
        E[loc[e]:loc[e + 1]] = err
    JTJ = J.T @ J
    JTE = J.T @ E
    x = np.linalg.solve(JTJ, JTE)          # conjungate(JTJ, JTE): CG on an SPD system -> its exact solution
    if dense_out:
        return x, J, E, JTJ, JTE
    return x, float(E @ E)


def ds_optimize(prob, back, cam_pose, params0, crit_type, max_count, eps, record=None):
    """The loop DoubleSideCalibration inherits (src/multicalib.cpp:462-514) on its own parameter vector."""
    params = np.asarray(params0, dtype=np.float64).copy()
    cam_pose = np.asarray(cam_pose, dtype=np.float64).reshape(-1, 6)
    change, it = 1.0, 0
    while True:
        if ((crit_type == 1 and it >= max_count) or (crit_type == 2 and change <= eps) or
                (crit_type == 3 and (change <= eps or it >= max_count))):
            break
        alpha = 0.95 ** (it + 1.0)
        x, cost = ds_compute_jacobian(prob, back, cam_pose, params)
        G = alpha * x
        params = params + G
        change = float(np.linalg.norm(G) / np.linalg.norm(params))
        if record is not None:
            record.append(dict(iter=it, cost_before=cost, change=change, params=params.copy()))
        it += 1
    return params, it, change


def ds_cost(prob, back, cam_pose, params):
    cam_pose = np.asarray(cam_pose, dtype=np.float64).reshape(-1, 6)
    nC = prob.n_camera
    c = 0.0
    for e in range(len(prob.edges)):
        cam, pv = prob.edges[e][0], prob.edges[e][1]
        row = pv - nC + 1
        _, _, err = ds_photo_doubleside_jacobian(prob, e, bool(back[e]), params[row * 6:row * 6 + 3], params[row * 6 + 3:row * 6 + 6],
                                                 cam_pose[cam, 0:3], cam_pose[cam, 3:6], params[0:3], params[3:6])
        c += float(err @ err)
    return c


def compute_project_error(prob, params, policy='fp64'):
    """src/multicalib.cpp:895-1006.  Returns dict with the reference's (quirky) mean error, the per-edge
    means, and the fp64 RMS of src/omnidir.cpp:1794-1802 (the north star's 'fp64 final RMS')."""
    f32 = policy == 'faithful_f32'
    p = np.asarray(params, dtype=np.float32 if f32 else np.float64)
    tot_err = np.float32(0) if f32 else 0.0
    tot_n = 0
    sq = 0.0
    npts = 0
    per_edge = []
    for e, (cam, pv, obj32, img32) in enumerate(prob.edges):
        dt = np.float32 if f32 else np.float64
        RP = cv2.Rodrigues(p[(pv - 1) * 6:(pv - 1) * 6 + 3].reshape(3, 1))[0].astype(dt)
        TP = p[(pv - 1) * 6 + 3:(pv - 1) * 6 + 6].reshape(3, 1).astype(dt)
        if cam == 0:
            R, T = RP, TP
        else:
            RC = cv2.Rodrigues(p[(cam - 1) * 6:(cam - 1) * 6 + 3].reshape(3, 1))[0].astype(dt)
            TC = p[(cam - 1) * 6 + 3:(cam - 1) * 6 + 6].reshape(3, 1).astype(dt)
            R = (RC @ RP).astype(dt)
            T = (RC @ TP + TC).astype(dt)
        rvec = cv2.Rodrigues(R)[0].astype(dt)
        obj = obj32.astype(dt)
        if prob.cam_type[cam] == PINHOLE:
            proj = cv2.projectPoints(obj.reshape(-1, 1, 3), rvec, T, prob.K[cam], prob.dist[cam].reshape(1, -1))[0]
            proj = proj.reshape(-1, 2)
        else:
            proj, _ = omnidir_project_points(obj, rvec.astype(np.float64), T.astype(np.float64), prob.K[cam],
                                             float(prob.xi[cam]), prob.dist[cam], want_jac=False)
            proj = proj.astype(dt)
        err = img32.astype(dt) - proj.astype(dt)
        nrm = np.sqrt(err[:, 0] * err[:, 0] + err[:, 1] * err[:, 1]).astype(dt)
        epi = dt(0)
        for v in nrm:
            epi = dt(epi + v)
        per_edge.append(float(epi / dt(err.shape[0])))
        tot_err = dt(tot_err + epi)
        # :983 error.total(): PINHOLE error is N x 2 single channel (2N), OMNIDIRECTIONAL N x 1 x 2ch (N)
        tot_n += err.shape[0] * (2 if prob.cam_type[cam] == PINHOLE else 1)
        e64 = img32.astype(np.float64) - proj.astype(np.float64)
        sq += float((e64 * e64).sum())
        npts += err.shape[0]
    return dict(mean_reproj_error=float(tot_err) / tot_n, per_edge=np.array(per_edge),
                rms=float(np.sqrt(sq / npts)), n_points=npts)


# --------------------------------------------------------------------------------------------
# omnidir::calibrate loop (row J)
# --------------------------------------------------------------------------------------------
def flags2idx(flags, n):
    """src/omnidir.cpp:2031-2076 (>= / subtract cascade, reproduced literally)."""
    idx = np.ones(6 * n + 10, dtype=np.int64)
    f = int(flags)
    if f >= CALIB_FIX_CENTER:
        idx[6 * n + 3] = 0; idx[6 * n + 4] = 0; f -= CALIB_FIX_CENTER
    if f >= CALIB_FIX_GAMMA:
        idx[6 * n] = 0; idx[6 * n + 1] = 0; f -= CALIB_FIX_GAMMA
    if f >= CALIB_FIX_XI:
        idx[6 * n + 5] = 0; f -= CALIB_FIX_XI
    if f >= CALIB_FIX_P2:
        idx[6 * n + 9] = 0; f -= CALIB_FIX_P2
    if f >= CALIB_FIX_P1:
        idx[6 * n + 8] = 0; f -= CALIB_FIX_P1
    if f >= CALIB_FIX_K2:
        idx[6 * n + 7] = 0; f -= CALIB_FIX_K2
    if f >= CALIB_FIX_K1:
        idx[6 * n + 6] = 0; f -= CALIB_FIX_K1
    if f >= CALIB_FIX_SKEW:
        idx[6 * n + 2] = 0
    return idx


def omni_compute_jacobian(obj_list, img_list, param, flags, epsilon):
    """src/omnidir.cpp:851-935.  param layout: [om_i,T_i]*n, fx, fy, s, cx, cy, xi, k1, k2, p1, p2."""
    n = len(obj_list)
    P = 6 * n + 10
    JTJ = np.zeros((P, P))
    JTE = np.zeros(P)
    K = np.array([[param[6 * n], param[6 * n + 2], param[6 * n + 3]],
                  [0, param[6 * n + 1], param[6 * n + 4]], [0, 0, 1.0]])
    D = param[6 * n + 6:6 * n + 10]
    xi = param[6 * n + 5]
    cost = 0.0
    for i in range(n):
        proj, jac = omnidir_project_points(obj_list[i], param[6 * i:6 * i + 3], param[6 * i + 3:6 * i + 6], K, xi, D)
        err = (np.asarray(img_list[i], dtype=np.float64).reshape(-1, 2) - proj).reshape(-1)
        JIn = jac[:, 6:16]
        JEx = jac[:, 0:6]
        JTJ[6 * n:, 6 * n:] += JIn.T @ JIn
        JTJ[6 * i:6 * i + 6, 6 * i:6 * i + 6] = JEx.T @ JEx
        JTJ[6 * i:6 * i + 6, 6 * n:] = JEx.T @ JIn
        JTJ[6 * n:, 6 * i:6 * i + 6] = JIn.T @ JEx
        JTE[6 * n:] += JIn.T @ err
        JTE[6 * i:6 * i + 6] = JEx.T @ err
        cost += float(err @ err)
    idx = flags2idx(flags, n).astype(bool)
    JTJs = JTJ[np.ix_(idx, idx)]
    JTEs = JTE[idx]
    JTJ_inv = np.linalg.inv(JTJs + epsilon)          # :934 -- scalar added to EVERY element
    return JTJ_inv, JTEs, idx, cost


def omni_calibrate_loop(obj_list, img_list, param0, flags, crit_type, max_count, eps, record=None):
    """src/omnidir.cpp:1119-1147."""
    cur = np.asarray(param0, dtype=np.float64).copy()
    change = 1.0
    it = 0
    while True:
        if ((crit_type == 1 and it >= max_count) or (crit_type == 2 and change <= eps) or
                (crit_type == 3 and (change <= eps or it >= max_count))):
            break
        alpha = 1 - (1 - 0.01) ** (it + 1.0)
        epsilon = 0.01 * 0.9 ** (it / 10.0)
        JTJ_inv, JTE, idx, cost = omni_compute_jacobian(obj_list, img_list, cur, flags, epsilon)
        Gs = alpha * (JTJ_inv @ JTE)
        G = np.zeros_like(cur)
        G[idx] = Gs                                   # fillFixed :2138-2153
        new = cur + G
        change = float(np.linalg.norm(G) / np.linalg.norm(cur))     # :1141 norm of the OLD parameters
        cur = new
        if record is not None:
            record.append(dict(iter=it, cost_before=cost, change=change, params=cur.copy()))
        it += 1
    return cur, it, change


def omni_rms(obj_list, img_list, param):
    """src/omnidir.cpp:1794-1802."""
    n = len(obj_list)
    K = np.array([[param[6 * n], param[6 * n + 2], param[6 * n + 3]],
                  [0, param[6 * n + 1], param[6 * n + 4]], [0, 0, 1.0]])
    D = param[6 * n + 6:6 * n + 10]
    xi = param[6 * n + 5]
    sq = 0.0
    cnt = 0
    for i in range(n):
        proj, _ = omnidir_project_points(obj_list[i], param[6 * i:6 * i + 3], param[6 * i + 3:6 * i + 6], K, xi, D,
                                         want_jac=False)
        e = np.asarray(img_list[i], dtype=np.float64).reshape(-1, 2) - proj
        sq += float((e * e).sum())
        cnt += e.shape[0]
    return float(np.sqrt(sq / cnt))


# --------------------------------------------------------------------------------------------
# omnidir::internal::initializeCalibration (src/omnidir.cpp:551-745) -- closed-form start of omnidir::calibrate,
# needed to re-enact the per-camera calibration MultiCameraCalibration::initialize runs before the rig loop
# (src/multicalib.cpp:276-279).  cv2 is used where the reference calls OpenCV (SVD, Rodrigues).
# --------------------------------------------------------------------------------------------
def _mean_repro_err(img, proj):
    d = np.asarray(img, dtype=np.float64).reshape(-1, 2) - np.asarray(proj, dtype=np.float64).reshape(-1, 2)
    return float(np.mean(np.sqrt((d * d).sum(axis=1))))          # computeMeanReproErr, src/omnidir.cpp:1571-1600


def omni_initialize_calibration(obj_list, img_list, size):
    """Returns (om list, t list, K, xi, idx) for the frames that survive the 100 px filter (:712-721)."""
    u0, v0 = size[0] // 2, size[1] // 2                             # integer division, :556-557
    n_img = len(img_list)
    om_all, t_all, gamma_all = [None] * n_img, [None] * n_img, [0.0] * n_img
    for ii in range(n_img):
        obj = np.asarray(obj_list[ii], dtype=np.float64).reshape(-1, 3)
        img = np.asarray(img_list[ii], dtype=np.float64).reshape(-1, 2)
        x, y = obj[:, 0], obj[:, 1]
        u, v = img[:, 0] - u0, img[:, 1] - v0
        rho2 = u * u + v * v
        M = np.stack([-v * x, -v * y, u * x, u * y, -v, u], axis=1)
        _, _, Vt = np.linalg.svd(M, full_matrices=True)
        V = Vt.T
        best = 1e5
        for coef in (1, -1):
            r11, r12, r21, r22, t1, t2 = (V[:6, 5] * coef)
            a, b = -(r11 * r12 + r21 * r22) ** 2, r11 * r11 + r21 * r21 - r12 * r12 - r22 * r22
            roots = np.roots([1.0, b, a])                           # solvePoly(a + b z + z^2)
            roots = sorted(roots.real, reverse=False)
            # the reference takes roots[0] if positive, else roots[1] (:613-616); cv::solvePoly's order is not
            # specified, the positive root is unique here (a <= 0)
            r31s = np.sqrt(max(roots))
            for coef2 in (1, -1):
                r31 = r31s * coef2
                r32 = -(r11 * r12 + r21 * r22) / r31
                r1 = np.array([r11, r21, r31]); r2 = np.array([r12, r22, r32]); t = np.array([t1, t2, 0.0])
                scale = 1.0 / np.linalg.norm(r1)
                r1, r2, t = r1 * scale, r2 * scale, t * scale
                n = x.size
                A = np.zeros((2 * n, 3))
                A[:n, 0] = (r1[1] * x + r2[1] * y + t[1]) / 2
                A[n:, 0] = (r1[0] * x + r2[0] * y + t[0]) / 2
                A[:n, 1] = -A[:n, 0] * rho2
                A[n:, 1] = -A[n:, 0] * rho2
                A[:n, 2] = -v
                A[n:, 2] = -u
                maxA = np.abs(A).max(axis=0)
                A = A / maxA
                B = np.concatenate([v * (r1[2] * x + r2[2] * y), u * (r1[2] * x + r2[2] * y)])
                res = np.linalg.pinv(A) @ B
                res = res / maxA
                with np.errstate(invalid="ignore"):
                    gamma = np.sqrt(res[0] / res[1])
                t[2] = res[2]
                r3 = np.cross(r1, r2)
                R = np.stack([r1, r2, r3], axis=1)
                om = cv2.Rodrigues(R)[0].ravel()
                if not np.isfinite(gamma):
                    continue
                Kc = np.array([[gamma, 0, u0], [0, gamma, v0], [0, 0, 1.0]])
                proj, _ = omnidir_project_points(obj, om, t, Kc, 1.0, np.zeros(4), want_jac=False)
                err = _mean_repro_err(img, proj)
                if err < best:
                    best, om_all[ii], t_all[ii], gamma_all[ii] = err, om, t.copy(), gamma
    g = sorted(gamma_all)
    gamma_final = g[len(g) // 2]                                    # nth_element at n/2, :705-707
    K = np.array([[gamma_final, 0, u0], [0, gamma_final, v0], [0, 0, 1.0]])
    idx, om_f, t_f = [], [], []
    for i in range(n_img):
        if om_all[i] is None:
            continue
        proj, _ = omnidir_project_points(np.asarray(obj_list[i], dtype=np.float64).reshape(-1, 3), om_all[i], t_all[i], K, 1.0,
                                         np.zeros(4), want_jac=False)
        if _mean_repro_err(img_list[i], proj) < 100:
            idx.append(i); om_f.append(om_all[i]); t_f.append(t_all[i])
    return om_f, t_f, K, 1.0, idx


# --------------------------------------------------------------------------------------------
# omnidir::stereoCalibrate: computeJacobianStereo src/omnidir.cpp:937-1020, the loop :1268-1296,
# flags2idxStereo :2078-2136, fillFixedStereo :2155-2170, initializeStereoCalibration :750-830, findMedian(3) :2172-2189,
# estimateUncertaintiesStereo :1804-1889.  Parameter vector (encodeParametersStereo :1570-1620):
#   [om, T (pose of camera 2 relative to camera 1) | om_i, T_i of the n frames in camera 1 | fx fy s cx cy xi k1 k2 p1 p2 of camera 1 | same of camera 2]
# --------------------------------------------------------------------------------------------
def flags2idx_stereo(flags, n):
    idx = np.ones(6 * (n + 1) + 20, dtype=np.int64)
    f = int(flags)
    o1, o2 = 6 * (n + 1), 6 * (n + 1) + 10
    for bit, cols in ((CALIB_FIX_CENTER, (3, 4)), (CALIB_FIX_GAMMA, (0, 1)), (CALIB_FIX_XI, (5,)), (CALIB_FIX_P2, (9,)),
                      (CALIB_FIX_P1, (8,)), (CALIB_FIX_K2, (7,)), (CALIB_FIX_K1, (6,))):
        if f >= bit:
            for c in cols:
                idx[o1 + c] = 0; idx[o2 + c] = 0
            f -= bit
    if f >= CALIB_FIX_SKEW:
        idx[o1 + 2] = 0; idx[o2 + 2] = 0
    return idx


def _stereo_unpack(param, n):
    o1, o2 = 6 * (n + 1), 6 * (n + 1) + 10
    K1 = np.array([[param[o1], param[o1 + 2], param[o1 + 3]], [0, param[o1 + 1], param[o1 + 4]], [0, 0, 1.0]])
    K2 = np.array([[param[o2], param[o2 + 2], param[o2 + 3]], [0, param[o2 + 1], param[o2 + 4]], [0, 0, 1.0]])
    return K1, param[o1 + 5], param[o1 + 6:o1 + 10], K2, param[o2 + 5], param[o2 + 6:o2 + 10]


def omni_stereo_jacobian(obj_list, img1_list, img2_list, param, flags, epsilon):
    """Dense J (rows: per frame the left image's 2N rows, then the right image's 2N rows) exactly as :937-1020."""
    n = len(obj_list)
    P = 6 * (n + 1) + 20
    K1, xi1, D1, K2, xi2, D2 = _stereo_unpack(param, n)
    om, T = param[0:3], param[3:6]
    rows, errs = [], []
    for i in range(n):
        obj = np.asarray(obj_list[i], dtype=np.float64).reshape(-1, 3)
        npt = obj.shape[0]
        om1, T1 = param[6 + 6 * i:9 + 6 * i], param[9 + 6 * i:12 + 6 * i]
        p1, j1 = omnidir_project_points(obj, om1, T1, K1, xi1, D1)
        J = np.zeros((4 * npt, P))
        J[:2 * npt, 6 * (n + 1):6 * (n + 1) + 10] = j1[:, 6:16]
        J[:2 * npt, 6 + 6 * i:12 + 6 * i] = j1[:, 0:6]
        e1 = (np.asarray(img1_list[i], dtype=np.float64).reshape(-1, 2) - p1).reshape(-1)
        (om2, T2, dom2dom1, dom2dT1, dom2dom, dom2dT, dT2dom1, dT2dT1, dT2dom, dT2dT) = compose_motion(om1, T1, om, T)
        p2, j2 = omnidir_project_points(obj, om2, T2, K2, xi2, D2)
        e2 = (np.asarray(img2_list[i], dtype=np.float64).reshape(-1, 2) - p2).reshape(-1)
        J[2 * npt:, 0:3] = j2[:, 0:3] @ dom2dom + j2[:, 3:6] @ dT2dom
        J[2 * npt:, 3:6] = j2[:, 0:3] @ dom2dT + j2[:, 3:6] @ dT2dT
        J[2 * npt:, 6 + 6 * i:9 + 6 * i] = j2[:, 0:3] @ dom2dom1 + j2[:, 3:6] @ dT2dom1
        J[2 * npt:, 9 + 6 * i:12 + 6 * i] = j2[:, 0:3] @ dom2dT1 + j2[:, 3:6] @ dT2dT1
        J[2 * npt:, 6 * (n + 1) + 10:6 * (n + 1) + 20] = j2[:, 6:16]
        rows.append(J); errs.append(np.concatenate([e1, e2]))
    J = np.vstack(rows); E = np.concatenate(errs)
    idx = flags2idx_stereo(flags, n).astype(bool)
    JTJ = (J.T @ J)[np.ix_(idx, idx)]
    JTE = (J.T @ E)[idx]
    return np.linalg.inv(JTJ + epsilon), JTE, idx, float(E @ E), E


def omni_stereo_calibrate_loop(obj_list, img1_list, img2_list, param0, flags, crit_type, max_count, eps, record=None):
    """src/omnidir.cpp:1268-1296 (same alpha / epsilon schedule as the single-camera loop)."""
    cur = np.asarray(param0, dtype=np.float64).copy()
    change, it = 1.0, 0
    while True:
        if ((crit_type == 1 and it >= max_count) or (crit_type == 2 and change <= eps) or
                (crit_type == 3 and (change <= eps or it >= max_count))):
            break
        alpha = 1 - (1 - 0.01) ** (it + 1.0)
        epsilon = 0.01 * 0.9 ** (it / 10.0)
        JTJ_inv, JTE, idx, cost, _ = omni_stereo_jacobian(obj_list, img1_list, img2_list, cur, flags, epsilon)
        G = np.zeros_like(cur)
        G[idx] = alpha * (JTJ_inv @ JTE)
        change = float(np.linalg.norm(G) / np.linalg.norm(cur))
        cur = cur + G
        if record is not None:
            record.append(dict(iter=it, cost_before=cost, change=change, params=cur.copy()))
        it += 1
    return cur, it, change


def omni_stereo_uncertainties(obj_list, img1_list, img2_list, param, flags):
    """estimateUncertaintiesStereo :1804-1889 -> (errors = 3 s sqrt(diag((J^T J)^-1)) over the free parameters, std_error (x, y), rms).
    The reference projects the right image through Rodrigues(R R1) (a log / exp round trip); the residuals are the same."""
    JTJ_inv, _, idx, _, E = omni_stereo_jacobian(obj_list, img1_list, img2_list, param, flags, 0.0)
    e = E.reshape(-1, 2)
    N = e.shape[0]
    std_error = e.std(axis=0) * np.sqrt(N / (N - 1.0))
    s = E.std() * np.sqrt(2.0 * N / (2.0 * N - 1.0))
    errors = 3 * s * np.sqrt(np.diag(JTJ_inv))
    rms = float(np.sqrt((e * e).sum() / N))
    return errors, std_error, rms, idx


def _find_median(v):
    """findMedian :2172-2181, with its even / odd quirk kept."""
    t = np.sort(np.asarray(v, dtype=np.float64))
    m = t.size
    return t[m // 2] if m % 2 == 0 else 0.5 * (t[m // 2] + t[m // 2 - 1])


def omni_initialize_stereo(obj_list, img1_list, img2_list, size1, size2, flags=0):
    """initializeStereoCalibration :750-830 -> (param0, idx of the frames both cameras kept)."""
    res = []
    for imgs, size in ((img1_list, size1), (img2_list, size2)):
        om, t, K, xi, idx = omni_initialize_calibration(obj_list, imgs, size)
        n = len(idx)
        p0 = np.concatenate([np.concatenate([np.concatenate([om[i], t[i]]) for i in range(n)]), [K[0, 0], K[1, 1], 0.0, K[0, 2], K[1, 2], xi, 0, 0, 0, 0]])
        p, _, _ = omni_calibrate_loop([obj_list[i] for i in idx], [imgs[i] for i in idx], p0, flags, 3, 100, 1e-6)
        res.append((idx, p, n))
    (idx1, p1, n1), (idx2, p2, n2) = res
    inter = [i for i in idx1 if i in idx2]
    om_est, t_est, omL, tL = [], [], [], []
    for fr in inter:
        a, b = idx1.index(fr), idx2.index(fr)
        R1 = cv2.Rodrigues(p1[6 * a:6 * a + 3])[0]; R2 = cv2.Rodrigues(p2[6 * b:6 * b + 3])[0]
        T1, T2 = p1[6 * a + 3:6 * a + 6], p2[6 * b + 3:6 * b + 6]
        RLR = R2 @ R1.T
        om_est.append(cv2.Rodrigues(RLR)[0].ravel()); t_est.append(T2 - RLR @ T1)
        omL.append(p1[6 * a:6 * a + 3]); tL.append(T1)
    om_est, t_est = np.array(om_est), np.array(t_est)
    om0 = np.array([_find_median(om_est[:, k]) for k in range(3)])
    t0 = np.array([_find_median(t_est[:, k]) for k in range(3)])
    n = len(inter)
    param0 = np.concatenate([om0, t0, np.concatenate([np.concatenate([omL[i], tL[i]]) for i in range(n)]), p1[6 * n1:6 * n1 + 10], p2[6 * n2:6 * n2 + 10]])
    return param0, inter
