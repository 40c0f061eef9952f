"""Generates the committed golden fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

Sources of truth:
  cv2_primitives.npz   OpenCV 4.13.0 via Python cv2: Rodrigues (+3x9 Jacobian), composeRT (all 10 outputs),
                       projectPoints (+ Jacobian columns 0..5) for 0/4/5/8 distortion coefficients      [tier 0]
  omni_points.npz      numpy transcription of /root/reference/src/omnidir.cpp:126-243 (oracle/dense_reenact.py),
                       i.e. cv::omnidir::projectPoints with its 2N x 16 Jacobian                        [tier 1]
  rig_dense.npz        dense literal re-enactment of optimizeExtrinsics (/root/reference/src/multicalib.cpp:462-514,
                       593-703, 717-824) on seeded toy rigs: iterates in both precision policies          [tier 2]
Nothing here reads /root/reference at run time; the reference itself cannot be built in this image.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

from oracle import dense_reenact as dr  # noqa: E402
from tests import rigs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def primitives():
    rng = np.random.default_rng(20261018)
    d = {}
    oms, Rs, Js = [], [], []
    for sc in [1e-7, 1e-3, 0.05, 0.3, 1.0, 2.0, 3.0]:
        for _ in range(3):
            om = rng.standard_normal(3); om *= sc / np.linalg.norm(om)
            R, J = cv2.Rodrigues(om)
            oms.append(om); Rs.append(R); Js.append(J)
    d["rod_om"] = np.array(oms); d["rod_R"] = np.array(Rs); d["rod_J"] = np.array(Js)
    ins, outs = [], []
    for _ in range(12):
        om1 = rng.standard_normal(3) * 0.6; t1 = rng.standard_normal(3) * 700
        om2 = rng.standard_normal(3) * 0.6; t2 = rng.standard_normal(3) * 300
        r = cv2.composeRT(om1, t1, om2, t2)
        ins.append(np.concatenate([om1, t1, om2, t2]))
        outs.append(np.concatenate([np.asarray(x).ravel() for x in r]))
    d["compose_in"] = np.array(ins); d["compose_out"] = np.array(outs)
    obj = rng.uniform(-250, 250, (30, 3)); obj[:, 2] *= 0.1
    om = np.array([0.21, -0.33, 0.12]); T = np.array([35.0, -60.0, 1400.0])
    K5 = np.array([1003.5, 998.25, 951.0, 546.5, 0.0])
    K = np.array([[K5[0], 0, K5[2]], [0, K5[1], K5[3]], [0, 0, 1.0]])
    dist8 = np.array([-0.21, 0.06, 1.1e-3, -0.7e-3, 0.012, 0.02, -0.01, 0.004])
    d["pin_obj"] = obj; d["pin_om"] = om; d["pin_T"] = T; d["pin_K5"] = K5; d["pin_dist8"] = dist8
    for nd in (0, 4, 5, 8):
        p, j = cv2.projectPoints(obj.reshape(-1, 1, 3), om, T, K, dist8[:nd] if nd else None)
        d["pin_proj_%d" % nd] = p.reshape(-1, 2); d["pin_jac_%d" % nd] = j[:, :6]
    # camodocal/PinholeCamera_test.cc known answers (same 4-coefficient radtan model)
    Kc = np.array([[712.557492, 0, 370.075592], [0, 714.825860, 244.759309], [0, 0, 1.0]])
    Dc = np.array([-0.473, 0.273, -0.001, 0.001])
    p, _ = cv2.projectPoints(np.array([[[0.0, 0, 1]], [[1.0, -1, 4]]]), np.zeros(3), np.zeros(3), Kc, Dc)
    d["camodocal_K5"] = np.array([712.557492, 714.825860, 370.075592, 244.759309, 0.0]); d["camodocal_D"] = Dc
    d["camodocal_uv"] = p.reshape(-1, 2)
    np.savez(os.path.join(OUT, "cv2_primitives.npz"), **d)


def omni():
    rng = np.random.default_rng(7)
    obj = rng.uniform(-300, 300, (25, 3)); obj[:, 2] *= 0.05
    om = np.array([-0.15, 0.4, 0.9]); T = np.array([-80.0, 40.0, 900.0])
    K5 = np.array([612.5, 640.25, 955.0, 530.0, 0.75]); xi = 1.25
    K = np.array([[K5[0], K5[4], K5[2]], [0, K5[1], K5[3]], [0, 0, 1.0]])
    D = np.array([-0.06, 0.012, 1.5e-3, -0.8e-3])
    p, j = dr.omnidir_project_points(obj, om, T, K, xi, D)
    np.savez(os.path.join(OUT, "omni_points.npz"), obj=obj, om=om, T=T, K5=K5, xi=xi, D=D, proj=p, jac=j)


def rig_dense():
    d = {}
    for name, kw in (("pin", dict(n_cam=3, n_frame=8, cam_models=[0, 0, 0], seed=41)),
                     ("omni", dict(n_cam=2, n_frame=7, cam_models=[1, 1], seed=42)),
                     ("mixed", dict(n_cam=3, n_frame=9, cam_models=[0, 1, 0], seed=43, views_per_frame=3, ragged=True))):
        rig = rigs.make_rig(**kw)
        P = rigs.to_dense_problem(rig)
        for pol in ("fp64", "faithful_f32"):
            rec = []
            p, it, ch = dr.optimize_extrinsics(P, rig["params_init"], 3, 12, 1e-7, pol, rec)
            d["%s_%s_iters" % (name, pol)] = it
            d["%s_%s_change" % (name, pol)] = ch
            d["%s_%s_params" % (name, pol)] = p
            d["%s_%s_trace" % (name, pol)] = np.array([[r["cost_before"], r["change"]] for r in rec])
            d["%s_%s_iterates" % (name, pol)] = np.array([r["params"] for r in rec])
            e = dr.compute_project_error(P, p, pol)
            d["%s_%s_err" % (name, pol)] = np.array([e["mean_reproj_error"], e["rms"], e["n_points"]])
            d["%s_%s_per_edge" % (name, pol)] = e["per_edge"]
        x, J, E, JTJ, JTE = dr.compute_jacobian_extrinsic(P, rig["params_init"], "fp64", dense_out=True)
        d[name + "_x0"] = x; d[name + "_JTE0"] = JTE; d[name + "_JTJ0_diag"] = np.diag(JTJ).copy()
        d[name + "_kw"] = np.array(repr(kw))
    np.savez_compressed(os.path.join(OUT, "rig_dense.npz"), **d)




def omni_fixture():
    """tutorials/data/omni_calib_data.xml (15 frames x 54 points, 1280x960; objectPoints '3d', imagePoints '2f') through
    the restated omnidir::calibrate loop.  The reference's closed-form initialisation (initializeCalibration) is out of
    scope, so the start is: intrinsics guess (f=400, centre = image centre, xi=1, D=0) + per-frame PnP on the rays
    unprojected with that guess.  Object points are rounded to float32 (the GPU path stores 20 B per corner)."""
    src = "/root/reference/tutorials/data/omni_calib_data.xml"
    if not os.path.exists(src):
        print("reference fixture not available, skipping omni_fixture")
        return
    fs = cv2.FileStorage(src, cv2.FILE_STORAGE_READ)
    on, im = fs.getNode("objectPoints"), fs.getNode("imagePoints")
    objs = [on.at(i).mat().reshape(-1, 3).astype(np.float32).astype(np.float64) for i in range(on.size())]
    imgs = [im.at(i).mat().reshape(-1, 2).astype(np.float32).astype(np.float64) for i in range(im.size())]
    n = len(objs)
    f0, cx, cy, xi = 400.0, 640.0, 480.0, 1.0
    poses = []
    for o, p in zip(objs, imgs):
        xu, yu = (p[:, 0] - cx) / f0, (p[:, 1] - cy) / f0
        r2 = xu * xu + yu * yu
        z = (xi + np.sqrt(1 + (1 - xi * xi) * r2)) / (r2 + 1)
        Xs = np.stack([z * xu, z * yu, z - xi], axis=1)
        npix = (Xs[:, :2] / Xs[:, 2:3]).reshape(-1, 1, 2)
        ok, rv, tv = cv2.solvePnP(o.reshape(-1, 1, 3), npix, np.eye(3), None)
        poses.append(np.concatenate([rv.ravel(), tv.ravel()]))
    p0 = np.concatenate([np.array(poses).ravel(), [f0, f0, 0.0, cx, cy, xi, 0, 0, 0, 0]])
    off = np.concatenate([[0], np.cumsum([o.shape[0] for o in objs])]).astype(np.int64)
    d = dict(off=off, obj=np.concatenate(objs).astype(np.float32), img=np.concatenate(imgs).astype(np.float32), p0=p0)
    for flags, crit in ((0, (3, 200, 1e-8)), (0, (3, 200, 1e-4)), (2, (3, 300, 1e-8))):
        rec = []
        p, it, ch = dr.omni_calibrate_loop(objs, imgs, p0, flags, crit[0], crit[1], crit[2], rec)
        key = "f%d_c%d_%d" % (flags, crit[0], crit[1]) + ("_e4" if crit[2] == 1e-4 else "")
        d[key + "_params"] = p; d[key + "_iters"] = it; d[key + "_change"] = ch
        d[key + "_rms"] = dr.omni_rms(objs, imgs, p)
        d[key + "_iter3"] = rec[2]["params"]
        print("omni fixture", key, "iters", it, "rms", d[key + "_rms"], "xi", p[6 * n + 5], "f", p[6 * n], p[6 * n + 1])
    np.savez_compressed(os.path.join(OUT, "omni_fixture.npz"), **d)


def stereo_rig_fixture():
    """BASELINE configs[0] substitute: the only real multi-camera data the reference ships,
    tutorials/data/omni_stereocalib_data.xml (39 frames x 48 corners, two 704x576 cameras), taken through the reference's
    own sequence -- per-camera omnidir::calibrate with its closed-form initialisation and criteria (COUNT+EPS, 300, 1e-7)
    (src/multicalib.cpp:276-279, src/omnidir.cpp:551-745, 1119-1147), spanning-tree pose chaining (:380-420), then the rig
    loop with the base-class default TermCriteria(COUNT, 20) (:462-514) -- by the literal dense re-enactment
    (oracle/dense_reenact.py, fp64 policy).  The survey's independent re-enactment reports: cam0 37/39 frames, RMS
    0.446242, xi 5.7313; cam1 36/39, RMS 0.406593, xi 1.5353; rig 40 vertices, 73 edges, RMS 0.540273 -> 0.456754,
    camera-1 tvec (-158.800, -19.662, -5.356)."""
    src = "/root/reference/tutorials/data/omni_stereocalib_data.xml"
    if not os.path.exists(src):
        print("reference fixture not available, skipping stereo_rig_fixture")
        return
    fs = cv2.FileStorage(src, cv2.FILE_STORAGE_READ)
    on = fs.getNode("objectPoints")
    objs = [on.at(i).mat().reshape(-1, 3) for i in range(on.size())]
    sz = fs.getNode("imageSize1")
    size = (int(sz.at(0).real()), int(sz.at(1).real()))
    d = {}
    cams = []
    for cam, key in ((0, "imagePoints1"), (1, "imagePoints2")):
        node = fs.getNode(key)
        imgs = [node.at(i).mat().reshape(-1, 2) for i in range(node.size())]
        om, t, K, xi, idx = dr.omni_initialize_calibration(objs, imgs, size)
        n = len(idx)
        p0 = np.concatenate([np.concatenate([np.concatenate([om[i], t[i]]) for i in range(n)]),
                             [K[0, 0], K[1, 1], 0.0, K[0, 2], K[1, 2], xi, 0, 0, 0, 0]])
        o_k, i_k = [objs[i] for i in idx], [imgs[i] for i in idx]
        p, it, ch = dr.omni_calibrate_loop(o_k, i_k, p0, 0, 3, 300, 1e-7)
        rms = dr.omni_rms(o_k, i_k, p)
        print("stereo fixture cam", cam, "kept", n, "iters", it, "rms", rms, "xi", p[6 * n + 5], "f", p[6 * n], p[6 * n + 1])
        off = np.concatenate([[0], np.cumsum([o.shape[0] for o in o_k])]).astype(np.int64)
        d["cam%d_idx" % cam] = np.array(idx, dtype=np.int32); d["cam%d_off" % cam] = off
        d["cam%d_obj" % cam] = np.concatenate(o_k).astype(np.float32); d["cam%d_img" % cam] = np.concatenate(i_k).astype(np.float32)
        d["cam%d_p0" % cam] = p0; d["cam%d_params" % cam] = p; d["cam%d_iters" % cam] = it; d["cam%d_rms" % cam] = rms
        cams.append(dict(idx=idx, p=p, n=n, imgs=imgs))
    # the rig: vertices 0, 1 = cameras; photo vertices in first-seen order (cameras outer loop, frames in index order);
    # the base class keeps every frame that survived a camera's own calibration (no multi-camera filter, src/multicalib.cpp:296-321)
    nC = 2
    ts2v, vts = {}, [-1, -1]
    edges, edge_T = [], []
    for c in range(nC):
        cc = cams[c]
        for j, frame in enumerate(cc["idx"]):
            if frame not in ts2v:
                ts2v[frame] = len(vts); vts.append(frame)
            pv = ts2v[frame]
            om, tt = cc["p"][6 * j:6 * j + 3], cc["p"][6 * j + 3:6 * j + 6]
            T = np.eye(4, dtype=np.float32)
            T[:3, :3] = cv2.Rodrigues(om.astype(np.float32))[0]; T[:3, 3] = tt.astype(np.float32)   # CV_32F transform, :300-318
            edges.append((c, pv, objs[frame].astype(np.float32), cc["imgs"][frame].astype(np.float32)))
            edge_T.append(T)
    nV = len(vts)
    # initialize(): BFS from vertex 0 over the camera-photo graph, neighbours in index order, last edge wins (:380-420)
    pose = [np.eye(4, dtype=np.float32) for _ in range(nV)]
    adj = [dict() for _ in range(nV)]
    for e, (c, pv, _, _) in enumerate(edges):
        adj[c][pv] = e; adj[pv][c] = e
    seen, order, pre, pre_e = {0}, [0], {}, {}
    q = [0]
    while q:
        v = q.pop(0)
        for w in sorted(adj[v]):
            if w not in seen:
                seen.add(w); pre[w] = v; pre_e[w] = adj[v][w]; q.append(w); order.append(w)
    for v in order[1:]:
        T = edge_T[pre_e[v]].astype(np.float64); P = pose[pre[v]].astype(np.float64)
        pose[v] = ((T @ np.linalg.inv(P)) if v < nC else (np.linalg.inv(P) @ T)).astype(np.float32)
    p_init = np.zeros(6 * (nV - 1), dtype=np.float32)
    for v in range(1, nV):
        p_init[6 * (v - 1):6 * (v - 1) + 3] = cv2.Rodrigues(pose[v][:3, :3].astype(np.float64))[0].ravel()
        p_init[6 * (v - 1) + 3:6 * v] = pose[v][:3, 3]
    K = np.zeros((nC, 3, 3)); xi = np.zeros(nC); dist = []
    for c in range(nC):
        pc, n = cams[c]["p"], cams[c]["n"]
        # intrinsics are kept CV_32F by the rig class (src/multicalib.cpp:281-283)
        f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
        K[c] = f32([[pc[6 * n], pc[6 * n + 2], pc[6 * n + 3]], [0, pc[6 * n + 1], pc[6 * n + 4]], [0, 0, 1]])
        xi[c] = f32(pc[6 * n + 5]); dist.append(f32(pc[6 * n + 6:6 * n + 10]))
    prob = dr.RigProblem([dr.OMNIDIRECTIONAL] * nC, K, dist, xi, edges, nV)
    rec = []
    p_fin, it, ch = dr.optimize_extrinsics(prob, p_init.astype(np.float64), 1, 20, 1e-7, policy="fp64", record=rec)
    err = dr.compute_project_error(prob, p_fin, policy="fp64")
    rms_seq = []
    for r in rec[:6]:
        rms_seq.append(dr.compute_project_error(prob, r["params"], policy="fp64")["rms"])
    rms0 = dr.compute_project_error(prob, p_init.astype(np.float64), policy="fp64")["rms"]
    print("stereo rig: vertices", nV, "edges", len(edges), "params", p_init.size, "corners", sum(e[2].shape[0] for e in edges))
    print("  rms init", rms0, "then", rms_seq, "final", err["rms"], "mean|e|", err["mean_reproj_error"] * 1.0)
    print("  camera 1 rvec", p_fin[0:3], "tvec", p_fin[3:6])
    d.update(rig_edge_cam=np.array([e[0] for e in edges], dtype=np.int32), rig_edge_pv=np.array([e[1] for e in edges], dtype=np.int32),
             rig_edge_off=np.concatenate([[0], np.cumsum([e[2].shape[0] for e in edges])]).astype(np.int64),
             rig_obj=np.concatenate([e[2] for e in edges]), rig_img=np.concatenate([e[3] for e in edges]),
             rig_vertex_timestamp=np.array(vts, dtype=np.int32), rig_K=K, rig_xi=xi, rig_dist=np.array(dist),
             rig_p_init=p_init.astype(np.float64), rig_p_final=p_fin, rig_iters=it, rig_rms_init=rms0, rig_rms_seq=np.array(rms_seq),
             rig_rms=err["rms"], rig_mean_error=err["mean_reproj_error"], rig_iter3=rec[2]["params"])
    np.savez_compressed(os.path.join(OUT, "stereo_rig_fixture.npz"), **d)


def stereo_ba_fixture():
    """cv::omnidir::stereoCalibrate on tutorials/data/omni_stereocalib_data.xml by the literal dense re-enactment
    (oracle/dense_reenact.py): initializeStereoCalibration (two omnidir::calibrate runs with criteria (3, 100, 1e-6), frame
    intersection, median relative pose), the loop with the tutorial's criteria (COUNT+EPS, 200, 1e-6), and
    estimateUncertaintiesStereo.  Golden vectors for SURVEY.md 8(f) row 3."""
    src = "/root/reference/tutorials/data/omni_stereocalib_data.xml"
    if not os.path.exists(src):
        print("reference fixture not available, skipping stereo_ba_fixture")
        return
    fs = cv2.FileStorage(src, cv2.FILE_STORAGE_READ)
    on, n1, n2 = fs.getNode("objectPoints"), fs.getNode("imagePoints1"), fs.getNode("imagePoints2")
    objs = [on.at(i).mat().reshape(-1, 3) for i in range(on.size())]
    im1 = [n1.at(i).mat().reshape(-1, 2) for i in range(n1.size())]
    im2 = [n2.at(i).mat().reshape(-1, 2) for i in range(n2.size())]
    s1 = fs.getNode("imageSize1"); s2 = fs.getNode("imageSize2")
    size1 = (int(s1.at(0).real()), int(s1.at(1).real())); size2 = (int(s2.at(0).real()), int(s2.at(1).real()))
    p0, inter = dr.omni_initialize_stereo(objs, im1, im2, size1, size2, 0)
    # the GPU path stores the points as float32 (20 B per corner): the golden run uses the same rounded data
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    o_k, a_k, b_k = [f32(objs[i]) for i in inter], [f32(im1[i]) for i in inter], [f32(im2[i]) for i in inter]
    d = dict(idx=np.array(inter, dtype=np.int32), off=np.concatenate([[0], np.cumsum([o.shape[0] for o in o_k])]).astype(np.int64),
             obj=np.concatenate(o_k).astype(np.float32), img1=np.concatenate(a_k).astype(np.float32),
             img2=np.concatenate(b_k).astype(np.float32), p0=p0)
    for flags, crit in ((0, (3, 200, 1e-6)), (2 + 64, (1, 30, 0.0))):
        rec = []
        p, it, ch = dr.omni_stereo_calibrate_loop(o_k, a_k, b_k, p0, flags, crit[0], crit[1], crit[2], rec)
        err, sd, rms, idx = dr.omni_stereo_uncertainties(o_k, a_k, b_k, p, flags)
        key = "f%d" % flags
        full = np.zeros(p.size); full[idx] = err
        d[key + "_params"] = p; d[key + "_iters"] = it; d[key + "_change"] = ch; d[key + "_iter3"] = rec[2]["params"]
        d[key + "_errors"] = full; d[key + "_std"] = sd; d[key + "_rms"] = rms
        n = len(inter)
        print("stereo BA flags", flags, "frames", n, "iters", it, "rms", rms, "om", p[:3], "T", p[3:6], "xi", p[6 * (n + 1) + 5], p[6 * (n + 1) + 15])
    np.savez_compressed(os.path.join(OUT, "stereo_ba_fixture.npz"), **d)


if __name__ == "__main__":
    primitives()
    omni()
    rig_dense()
    omni_fixture()
    stereo_rig_fixture()
    stereo_ba_fixture()
    print("golden fixtures written to", OUT)
