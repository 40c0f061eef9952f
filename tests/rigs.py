"""Small seeded synthetic rigs for the parity tests (numpy only; independent of both the oracle and the product's
own C++ generator).  Conventions follow SURVEY.md section 8(d): 9x6 board, 40 mm pitch, units mm, camera 0 =
identity, X_cam = R_cam (R_photo X + t_photo) + t_cam, observations stored float32, edge order = cameras outer
loop / timestamps inner loop, photo vertices in first-seen order (oracle/indexing.py)."""
from __future__ import annotations

import numpy as np

PINHOLE, OMNIDIR = 0, 1


def rodrigues(om):
    om = np.asarray(om, dtype=np.float64)
    th = np.linalg.norm(om)
    K = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def log_so3(R):
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-9:
        return v / 2
    return v * th / (2 * np.sin(th))


def project(model, K5, dist8, xi, Xc):
    """fp64 forward model for generating observations.  Xc (N,3)."""
    fx, fy, cx, cy, s = K5
    k1, k2, p1, p2, k3, k4, k5, k6 = dist8
    if model == PINHOLE:
        x = Xc[:, 0] / Xc[:, 2]; y = Xc[:, 1] / Xc[:, 2]
        r2 = x * x + y * y
        rad = (1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3) / (1 + k4 * r2 + k5 * r2 ** 2 + k6 * r2 ** 3)
        xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        return np.stack([fx * xd + cx, fy * yd + cy], axis=1)
    n = np.linalg.norm(Xc, axis=1)
    Xs = Xc / n[:, None]
    x = Xs[:, 0] / (Xs[:, 2] + xi); y = Xs[:, 1] / (Xs[:, 2] + xi)
    r2 = x * x + y * y
    rad = 1 + k1 * r2 + k2 * r2 ** 2
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([fx * xd + s * yd + cx, fy * yd + cy], axis=1)


def board(nx=9, ny=6, pitch=40.0):
    ii, jj = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    return np.stack([jj.ravel() * pitch, ii.ravel() * pitch, np.zeros(nx * ny)], axis=1)


def make_rig(n_cam=3, n_frame=12, cam_models=None, views_per_frame=2, seed=7, noise_px=0.3, init_rot=0.02,
             init_trans=10.0, ndist=5, nx=9, ny=6, ragged=False):
    """Returns a dict of arrays in the C-ABI layout plus params_true / params_init (float32-representable)."""
    rng = np.random.default_rng(seed)
    if cam_models is None:
        cam_models = [PINHOLE] * n_cam
    cam_models = list(cam_models)
    B = board(nx, ny)
    Bc = B.mean(axis=0)
    # cameras
    cam_R, cam_t, K5s, d8s, xis = [], [], [], [], []
    for c in range(n_cam):
        if c == 0:
            R = np.eye(3); t = np.zeros(3)
        else:
            th = np.deg2rad(-30 + 60.0 * c / max(n_cam - 1, 1))
            om = np.array([0, th, 0]) + np.deg2rad(2.0) * rng.standard_normal(3)
            R = rodrigues(om)
            centre = np.array([400 * np.sin(-th), 0, 400 * (1 - np.cos(th))]) + 5 * rng.standard_normal(3)
            t = -R @ centre
        cam_R.append(R); cam_t.append(t)
        d8 = np.zeros(8)
        d8[:4] = [0.05 * rng.standard_normal(), 0.01 * rng.standard_normal(), 1e-3 * rng.standard_normal(),
                  1e-3 * rng.standard_normal()]
        if cam_models[c] == PINHOLE:
            K5s.append([rng.uniform(950, 1050), rng.uniform(950, 1050), 960 + rng.uniform(-20, 20),
                        540 + rng.uniform(-20, 20), 0.0])
            if ndist >= 5:
                d8[4] = 0.002 * rng.standard_normal()
            if ndist >= 8:
                d8[5:8] = 0.01 * rng.standard_normal(3)
            xis.append(0.0)
        else:
            K5s.append([rng.uniform(500, 700), rng.uniform(500, 700), 960 + rng.uniform(-20, 20),
                        540 + rng.uniform(-20, 20), 0.5 * rng.standard_normal()])
            xis.append(rng.uniform(0.8, 1.5))
        d8s.append(d8)
    K5s = np.array(K5s, dtype=np.float32).astype(np.float64)   # intrinsics are stored CV_32F in the reference
    d8s = np.array(d8s, dtype=np.float32).astype(np.float64)
    xis = np.array(xis, dtype=np.float32).astype(np.float64)
    # frames: frame k seen by cameras k%nC, (k+1)%nC, ...
    frame_R, frame_t, views = [], [], []
    for k in range(n_frame):
        cams = sorted({(k + j) % n_cam for j in range(min(views_per_frame, n_cam))})
        for _ in range(200):
            axis = rng.standard_normal(3); axis /= np.linalg.norm(axis)
            tilt = rodrigues(axis * rng.uniform(0, np.deg2rad(25)))
            inplane = rodrigues(np.array([0, 0, rng.uniform(0, 2 * np.pi)]))
            Rb = tilt @ inplane
            # place the board centre in front of the first viewing camera, expressed in the reference frame
            c0 = cams[0]
            pc = np.array([rng.uniform(-150, 150), rng.uniform(-100, 100), rng.uniform(1200, 2000)])
            pw = cam_R[c0].T @ (pc - cam_t[c0])
            # face roughly towards the bisector of the viewing cameras
            Rw = cam_R[c0].T @ Rb
            tw = pw - Rw @ Bc
            ok = True
            for c in cams:
                Xc = (cam_R[c] @ (Rw @ B.T + tw[:, None]) + cam_t[c][:, None]).T
                if Xc[:, 2].min() < 300:
                    ok = False; break
                uv = project(cam_models[c], K5s[c], d8s[c], xis[c], Xc)
                if uv[:, 0].min() < 20 or uv[:, 0].max() > 1900 or uv[:, 1].min() < 20 or uv[:, 1].max() > 1060:
                    ok = False; break
            if ok and 300 < np.linalg.norm(tw) < 3000:
                break
        else:
            raise RuntimeError("could not place frame %d" % k)
        frame_R.append(Rw); frame_t.append(tw); views.append(cams)
    # edges in the reference order: cameras outer, timestamps inner; photo vertices first-seen
    pv_of_frame = {}
    edge_cam, edge_pv, edge_frame = [], [], []
    for c in range(n_cam):
        for k in range(n_frame):
            if c in views[k]:
                if k not in pv_of_frame:
                    pv_of_frame[k] = n_cam + len(pv_of_frame)
                edge_cam.append(c); edge_pv.append(pv_of_frame[k]); edge_frame.append(k)
    frame_of_pv = {v: k for k, v in pv_of_frame.items()}
    objs, imgs, off = [], [], [0]
    for c, k in zip(edge_cam, edge_frame):
        if ragged:
            keep = np.sort(rng.choice(B.shape[0], size=int(rng.integers(20, B.shape[0] + 1)), replace=False))
        else:
            keep = np.arange(B.shape[0])
        X = B[keep]
        Xc = (cam_R[c] @ (frame_R[k] @ X.T + frame_t[k][:, None]) + cam_t[c][:, None]).T
        uv = project(cam_models[c], K5s[c], d8s[c], xis[c], Xc) + noise_px * rng.standard_normal((X.shape[0], 2))
        objs.append(X.astype(np.float32)); imgs.append(uv.astype(np.float32))
        off.append(off[-1] + X.shape[0])
    n_vertex = n_cam + n_frame
    p_true = np.zeros(6 * (n_vertex - 1))
    for c in range(1, n_cam):
        p_true[6 * (c - 1):6 * (c - 1) + 3] = log_so3(cam_R[c]); p_true[6 * (c - 1) + 3:6 * c] = cam_t[c]
    for v in range(n_cam, n_vertex):
        k = frame_of_pv[v]
        p_true[6 * (v - 1):6 * (v - 1) + 3] = log_so3(frame_R[k]); p_true[6 * (v - 1) + 3:6 * v] = frame_t[k]
    p_init = p_true.copy()
    pert = rng.standard_normal(p_init.size)
    for v in range(1, n_vertex):
        p_init[6 * (v - 1):6 * (v - 1) + 3] += init_rot * pert[6 * (v - 1):6 * (v - 1) + 3]
        p_init[6 * (v - 1) + 3:6 * v] += init_trans * pert[6 * (v - 1) + 3:6 * v]
    p_init = p_init.astype(np.float32).astype(np.float64)          # buildParas stores CV_32F
    nd = np.array([ndist if m == PINHOLE else 4 for m in cam_models], dtype=np.int32)
    return dict(n_cam=n_cam, n_frame=n_frame, edge_cam=np.array(edge_cam, dtype=np.int32),
                edge_pv=np.array(edge_pv, dtype=np.int32), edge_off=np.array(off, dtype=np.int64),
                obj=np.concatenate(objs).astype(np.float32), img=np.concatenate(imgs).astype(np.float32),
                cam_model=np.array(cam_models, dtype=np.int32), cam_K5=K5s, cam_dist8=d8s, cam_ndist=nd,
                cam_xi=xis, params_true=p_true, params_init=p_init,
                timestamps=np.array([frame_of_pv[v] for v in range(n_cam, n_vertex)], dtype=np.int32))


def to_dense_problem(rig):
    """Convert to oracle.dense_reenact.RigProblem."""
    from oracle import dense_reenact as dr
    nC = rig["n_cam"]
    K = np.zeros((nC, 3, 3))
    for c in range(nC):
        fx, fy, cx, cy, s = rig["cam_K5"][c]
        K[c] = [[fx, s, cx], [0, fy, cy], [0, 0, 1]]
    dist = [rig["cam_dist8"][c][:rig["cam_ndist"][c]] for c in range(nC)]
    edges = []
    for e in range(rig["edge_cam"].size):
        a, b = rig["edge_off"][e], rig["edge_off"][e + 1]
        edges.append((int(rig["edge_cam"][e]), int(rig["edge_pv"][e]), rig["obj"][a:b], rig["img"][a:b]))
    return dr.RigProblem(rig["cam_model"], K, dist, rig["cam_xi"], edges, nC + rig["n_frame"])


def to_oracle_rig(rig):
    from oracle import oracle as orc
    return orc.Rig(rig["n_cam"], rig["n_frame"], rig["edge_cam"], rig["edge_pv"], rig["edge_off"], rig["obj"],
                   rig["img"], rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"])
