"""Deterministic synthetic rigs of the shapes BASELINE.json names, vectorised numpy.

Data synthesis only (no solver arithmetic): produces the arrays the C ABI takes.  Conventions:
  * 9x6 board, 40 mm pitch, X = (j*40, i*40, 0), row-major i in [0,6), j in [0,9); units mm so every |t| is in
    (300, 3000) (the reference asserts that, src/multicalib.cpp:107-113);
  * camera 0 = identity; camera c rotated about y by -30deg + 60deg*c/(nC-1) (+ N(0,2deg) per axis), centre on a
    400 mm arc; X_cam = R_cam (R_photo X + t_photo) + t_cam;
  * frame k is seen by cameras k mod nC, (k+1) mod nC, ... (views_per_frame of them); all corners of every view
    inside the 1920x1080 image with a 20 px margin (redrawn until they are);
  * observations = exact fp64 projection + N(0, noise_px), rounded to float32; intrinsics and the initial guess are
    float32-representable (the reference stores them CV_32F);
  * edge order = cameras outer loop, timestamps inner loop; photo vertices in first-seen order
    (src/mymulticalib.cpp:360-403, src/multicalib.cpp:323-346).
Counter-based RNG (Philox) keyed by (seed, stream) so that the cameras of a rig can be regenerated identically on
every rank while the frames differ per rank.
"""
from __future__ import annotations

import numpy as np

PINHOLE, OMNIDIRECTIONAL = 0, 1
IMG_W, IMG_H, MARGIN = 1920.0, 1080.0, 20.0


def _rng(seed, stream):
    return np.random.Generator(np.random.Philox(key=[int(seed), int(stream)]))


def rodrigues_batch(om):
    """om (...,3) -> R (...,3,3)"""
    om = np.asarray(om, dtype=np.float64)
    th = np.linalg.norm(om, axis=-1)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    a = np.where(small, 1.0 - th ** 2 / 6, np.sin(ths) / ths)
    b = np.where(small, 0.5 - th ** 2 / 24, (1 - np.cos(ths)) / ths ** 2)
    x, y, z = om[..., 0], om[..., 1], om[..., 2]
    K = np.zeros(om.shape[:-1] + (3, 3))
    K[..., 0, 1] = -z; K[..., 0, 2] = y; K[..., 1, 0] = z; K[..., 1, 2] = -x; K[..., 2, 0] = -y; K[..., 2, 1] = x
    K2 = K @ K
    return np.eye(3) + a[..., None, None] * K + b[..., None, None] * K2


def log_so3_batch(R):
    """R (...,3,3) -> om (...,3), rotation angle < pi assumed (synthetic rigs stay far from pi)."""
    c = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1) / 2, -1, 1)
    th = np.arccos(c)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], axis=-1)
    s = np.sin(th)
    k = np.where(th < 1e-9, 0.5, th / (2 * np.where(th < 1e-9, 1.0, s)))
    return v * k[..., None]


def project(model, K5, dist8, xi, Xc):
    """Forward camera models in fp64 (pinhole radtan/rational as cv::projectPoints; Mei as src/omnidir.cpp:146-165).
    model/K5/dist8/xi broadcast against Xc (...,3)."""
    fx, fy, cx, cy, s = [K5[..., i] for i in range(5)]
    k1, k2, p1, p2, k3, k4, k5, k6 = [dist8[..., i] for i in range(8)]
    X, Y, Z = Xc[..., 0], Xc[..., 1], Xc[..., 2]
    omni = np.asarray(model) == OMNIDIRECTIONAL
    n = np.sqrt(X * X + Y * Y + Z * Z)
    den = np.where(omni, Z / n + xi, Z)
    num_scale = np.where(omni, 1.0 / n, 1.0)
    x = X * num_scale / den
    y = Y * num_scale / den
    r2 = x * x + y * y
    rad = (1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3) / (1 + k4 * r2 + k5 * r2 ** 2 + k6 * r2 ** 3)
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    u = fx * xd + np.where(omni, s, 0.0) * yd + cx
    v = fy * yd + cy
    return np.stack([u, v], axis=-1)


def board(nx=9, ny=6, pitch=40.0):
    ii, jj = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    return np.stack([jj.ravel() * pitch, ii.ravel() * pitch, np.zeros(nx * ny)], axis=1)


def make_cameras(n_cam, seed, models=None, ndist=5):
    """Camera poses (truth) and float32-representable intrinsics.  models: list of PINHOLE/OMNIDIRECTIONAL."""
    rng = _rng(seed, 1)
    if models is None:
        models = [PINHOLE] * n_cam
    models = np.asarray(models, dtype=np.int32)
    th = np.deg2rad(-30.0 + 60.0 * np.arange(n_cam) / max(n_cam - 1, 1))
    om = np.zeros((n_cam, 3))
    om[:, 1] = th
    om += np.deg2rad(2.0) * rng.standard_normal((n_cam, 3))
    centre = np.stack([400 * np.sin(-th), np.zeros(n_cam), 400 * (1 - np.cos(th))], axis=1) + 5 * rng.standard_normal((n_cam, 3))
    om[0] = 0
    centre[0] = 0
    R = rodrigues_batch(om)
    t = -np.einsum("cij,cj->ci", R, centre)
    K5 = np.zeros((n_cam, 5)); d8 = np.zeros((n_cam, 8)); xi = np.zeros(n_cam)
    u = rng.uniform(size=(n_cam, 6)); g = rng.standard_normal((n_cam, 8))
    pin = models == PINHOLE
    K5[:, 0] = np.where(pin, 950 + 100 * u[:, 0], 500 + 200 * u[:, 0])
    K5[:, 1] = np.where(pin, 950 + 100 * u[:, 1], 500 + 200 * u[:, 1])
    K5[:, 2] = 960 + 40 * (u[:, 2] - 0.5)
    K5[:, 3] = 540 + 40 * (u[:, 3] - 0.5)
    K5[:, 4] = np.where(pin, 0.0, 0.5 * g[:, 7])
    xi = np.where(pin, 0.0, 0.8 + 0.7 * u[:, 4])
    d8[:, 0] = 0.05 * g[:, 0]; d8[:, 1] = 0.01 * g[:, 1]; d8[:, 2] = 1e-3 * g[:, 2]; d8[:, 3] = 1e-3 * g[:, 3]
    if ndist >= 5:
        d8[:, 4] = np.where(pin, 0.0, 0.0)       # k3 = 0 as in SURVEY.md 8(d)
    nd = np.where(pin, ndist, 4).astype(np.int32)
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    return dict(n_cam=n_cam, cam_model=models, cam_K5=f32(K5), cam_dist8=f32(d8), cam_ndist=nd, cam_xi=f32(xi),
                cam_R=R, cam_t=t, cam_om=om)


def make_rig(n_cam=8, n_frame=1000, seed=1002, models=None, views_per_frame=2, noise_px=0.3, init_rot=0.02,
             init_trans=10.0, ndist=5, frame_stream=0, cameras=None, perturb_cameras=True, board_distance=(1200.0, 2000.0),
             tilt_max_deg=30.0, lateral=120.0, min_depth=300.0, board_shape=(9, 6, 40.0)):
    """Full rig in the C-ABI layout.  frame_stream selects an independent set of frames for the same cameras (used to
    give every rank its own shard in the weak-scaling benchmark)."""
    cams = cameras if cameras is not None else make_cameras(n_cam, seed, models, ndist)
    n_cam = cams["n_cam"]
    rng = _rng(seed, 1000 + frame_stream)
    B = board(*board_shape)
    nB = B.shape[0]
    Bc = B.mean(axis=0)
    V = min(views_per_frame, n_cam)
    k = np.arange(n_frame)
    view_cam = (k[:, None] + np.arange(V)[None, :]) % n_cam           # (F, V)
    cR, ct = cams["cam_R"], cams["cam_t"]
    centre = -np.einsum("cji,cj->ci", cR, ct)                          # camera centres in the reference frame
    axis_w = cR[:, 2, :]                                               # optical axes in the reference frame
    frame_R = np.zeros((n_frame, 3, 3)); frame_t = np.zeros((n_frame, 3))
    todo = np.arange(n_frame)
    for _ in range(400):
        if todo.size == 0:
            break
        m = todo.size
        vc = view_cam[todo]
        bis = axis_w[vc].sum(axis=1); bis /= np.linalg.norm(bis, axis=1, keepdims=True)
        mid = centre[vc].mean(axis=1)
        # orthonormal frame around the bisector for lateral jitter
        up = np.tile(np.array([0.0, 1.0, 0.0]), (m, 1))
        ex = np.cross(up, bis); ex /= np.linalg.norm(ex, axis=1, keepdims=True)
        ey = np.cross(bis, ex)
        dist = rng.uniform(board_distance[0], board_distance[1], m)
        lat = rng.uniform(-lateral, lateral, (m, 2))
        pos = mid + bis * dist[:, None] + ex * lat[:, :1] + ey * lat[:, 1:]
        ax = rng.standard_normal((m, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
        tilt = rodrigues_batch(ax * rng.uniform(0, np.deg2rad(tilt_max_deg), m)[:, None])
        inpl = rodrigues_batch(np.stack([np.zeros(m), np.zeros(m), rng.uniform(0, 2 * np.pi, m)], axis=1))
        # board z axis facing back towards the cameras: base orientation = frame of the bisector rotated by pi about x
        base = np.stack([ex, -ey, -bis], axis=2)                       # columns: board x, y, z in the reference frame
        Rw = base @ tilt @ inpl
        tw = pos - np.einsum("fij,j->fi", Rw, Bc)
        Xw = np.einsum("fij,nj->fni", Rw, B) + tw[:, None, :]          # (m, nB, 3)
        ok = (np.linalg.norm(tw, axis=1) > min_depth) & (np.linalg.norm(tw, axis=1) < 3000)
        for v in range(V):
            c = vc[:, v]
            Xc = np.einsum("fij,fnj->fni", cR[c], Xw) + ct[c][:, None, :]
            uv = project(cams["cam_model"][c][:, None], cams["cam_K5"][c][:, None, :], cams["cam_dist8"][c][:, None, :],
                         cams["cam_xi"][c][:, None], Xc)
            T3 = np.einsum("fij,fj->fi", cR[c], tw) + ct[c]
            ok &= (Xc[..., 2].min(axis=1) > 0.5 * min_depth) & (np.linalg.norm(T3, axis=1) < 3000)
            ok &= (uv[..., 0].min(axis=1) > MARGIN) & (uv[..., 0].max(axis=1) < IMG_W - MARGIN)
            ok &= (uv[..., 1].min(axis=1) > MARGIN) & (uv[..., 1].max(axis=1) < IMG_H - MARGIN)
        sel = todo[ok]
        frame_R[sel] = Rw[ok]; frame_t[sel] = tw[ok]
        todo = todo[~ok]
    if todo.size:
        raise RuntimeError("could not place %d frames" % todo.size)
    # indexing: edges sorted by (camera, timestamp); photo vertices in first-seen order
    ev_cam = view_cam.ravel()
    ev_k = np.repeat(k, V)
    order = np.lexsort((ev_k, ev_cam))
    edge_cam = ev_cam[order].astype(np.int32)
    edge_k = ev_k[order]
    first_cam = view_cam.min(axis=1)
    vorder = np.lexsort((k, first_cam))                                # frames in first-seen order
    pv_of_frame = np.empty(n_frame, dtype=np.int64)
    pv_of_frame[vorder] = n_cam + np.arange(n_frame)
    edge_pv = pv_of_frame[edge_k].astype(np.int32)
    E = edge_cam.size
    edge_off = (np.arange(E + 1, dtype=np.int64) * nB)
    # observations, generated edge by edge in chunks to bound memory
    obj = np.empty((E * nB, 3), dtype=np.float32)
    img = np.empty((E * nB, 2), dtype=np.float32)
    obj[:] = np.tile(B.astype(np.float32), (E, 1))
    nrng = _rng(seed, 2000 + frame_stream)
    chunk = 20000
    for a in range(0, E, chunk):
        b = min(E, a + chunk)
        c = edge_cam[a:b]; kk = edge_k[a:b]
        Xw = np.einsum("fij,nj->fni", frame_R[kk], B) + frame_t[kk][:, None, :]
        Xc = np.einsum("fij,fnj->fni", cR[c], Xw) + ct[c][:, None, :]
        uv = project(cams["cam_model"][c][:, None], cams["cam_K5"][c][:, None, :], cams["cam_dist8"][c][:, None, :],
                     cams["cam_xi"][c][:, None], Xc)
        uv = uv + noise_px * nrng.standard_normal(uv.shape)
        img[a * nB:b * nB] = uv.reshape(-1, 2).astype(np.float32)
    # parameters
    n_vertex = n_cam + n_frame
    p_true = np.zeros((n_vertex - 1, 6))
    p_true[:n_cam - 1, :3] = log_so3_batch(cR[1:]); p_true[:n_cam - 1, 3:] = ct[1:]
    fr_of_pv = vorder                                                  # vertex n_cam + i  <->  frame vorder[i]
    p_true[n_cam - 1:, :3] = log_so3_batch(frame_R[fr_of_pv]); p_true[n_cam - 1:, 3:] = frame_t[fr_of_pv]
    prng = _rng(seed, 3000 + frame_stream)
    pert = prng.standard_normal(p_true.shape)
    crng = _rng(seed, 3)                                               # camera perturbation identical on every rank
    pert[:n_cam - 1] = crng.standard_normal((n_cam - 1, 6)) if perturb_cameras else 0.0
    p_init = p_true.copy()
    p_init[:, :3] += init_rot * pert[:, :3]
    p_init[:, 3:] += init_trans * pert[:, 3:]
    p_init = p_init.astype(np.float32).astype(np.float64)              # buildParas stores CV_32F
    return dict(n_cam=n_cam, n_frame=n_frame, edge_cam=edge_cam, edge_pv=edge_pv, edge_off=edge_off, obj=obj, img=img,
                cam_model=cams["cam_model"], cam_K5=cams["cam_K5"], cam_dist8=cams["cam_dist8"],
                cam_ndist=cams["cam_ndist"], cam_xi=cams["cam_xi"], params_true=p_true.ravel(),
                params_init=p_init.ravel(), timestamps=fr_of_pv.astype(np.int32), n_points=int(E * nB))


# BASELINE.json configs
CONFIGS = {
    2: dict(n_cam=8, n_frame=1000, seed=1002),
    # single Mei camera: boards close and strongly tilted, otherwise focal length and xi are not separately observable
    3: dict(n_cam=1, n_frame=5000, seed=1003, models="omni", board_distance=(250.0, 600.0), tilt_max_deg=50.0,
            lateral=250.0, min_depth=150.0),
    4: dict(n_cam=16, n_frame=10000, seed=1004, models="mixed"),
    5: dict(n_cam=64, n_frame=100000, seed=1005),
}


def make_double_side_rig(n_frame=60, seed=4001, theta=2.0, noise_px=0.3, init_rot=0.02, init_trans=10.0, third_every=3,
                         back_shape=(9, 6, 40.0)):
    """Synthetic problem of the double-sided board calibration (src/doubleSide.cpp): three FIXED pinhole cameras --
    camera 0 at the origin looks at the front pattern, camera 1 sits `theta` rad around the board and looks at the back
    pattern, camera 2 (0.35 rad off camera 0) sees the front pattern of every `third_every`-th frame -- and a board whose
    back pattern is related to the front pattern by D (X_front = R_D X_back + t_D).  theta stays well away from pi so
    that the reference's Rodrigues round trips (compose_motion) are accurate in the oracle.
    Returns the rig in the C-ABI layout plus edge_back, cam_pose (nC x 6), ds_params_true / ds_params_init
    ([D | frame poses] in photo-vertex order).  back_shape: the back pattern's own corner grid (the reference tells the
    sides apart by their corner count)."""
    cams = make_cameras(3, seed, [0, 0, 0], 5)
    B = board(); Bc = B.mean(axis=0)
    Bb = board(*back_shape)
    rng = _rng(seed, 50)
    centre_w = np.array([0.0, 0.0, 1500.0])

    def cam_at(angle):     # world -> camera for a camera on the circle around the board centre, looking at it
        R = rodrigues_batch(np.array([[0.0, angle, 0.0]]))[0]
        c = centre_w + 1500.0 * np.array([np.sin(angle), 0.0, -np.cos(angle)])
        return R, -R @ c
    poses = [cam_at(0.0), cam_at(theta), cam_at(-0.35)]
    cR = np.stack([p[0] for p in poses]); ct = np.stack([p[1] for p in poses])
    cam_pose = np.concatenate([log_so3_batch(cR), ct], axis=1)
    # D: the back pattern faces camera 1 the way the front pattern faces camera 0
    t0 = centre_w - Bc
    Rd = cR[1].T
    td = cR[1].T @ (t0 - ct[1]) - t0
    D_true = np.concatenate([log_so3_batch(Rd[None])[0], td])
    frame_R = np.zeros((n_frame, 3, 3)); frame_t = np.zeros((n_frame, 3))
    for f in range(n_frame):
        for _ in range(200):
            ax = rng.standard_normal(3); ax /= np.linalg.norm(ax)
            Rp = rodrigues_batch((ax * rng.uniform(0, 0.25))[None])[0]
            tp = t0 + rng.uniform(-60, 60, 3) + (Bc - Rp @ Bc)
            ok = True
            for c, back in ((0, False), (1, True), (2, False)):
                X = Bb @ Rd.T + td if back else B
                Xc = (X @ Rp.T + tp) @ cR[c].T + ct[c]
                uv = project(cams["cam_model"][c], cams["cam_K5"][c], cams["cam_dist8"][c], cams["cam_xi"][c], Xc)
                ok &= bool(Xc[:, 2].min() > 300 and uv[:, 0].min() > MARGIN and uv[:, 0].max() < IMG_W - MARGIN and
                           uv[:, 1].min() > MARGIN and uv[:, 1].max() < IMG_H - MARGIN)
            if ok:
                break
        else:
            raise RuntimeError("could not place frame %d" % f)
        frame_R[f] = Rp; frame_t[f] = tp
    # edges sorted by (camera, frame) as the reference loads them; photo vertices in first-seen order = frame order
    ev = [(0, f, 0) for f in range(n_frame)] + [(1, f, 1) for f in range(n_frame)] + \
         [(2, f, 0) for f in range(n_frame) if f % third_every == 0]
    edge_cam = np.array([e[0] for e in ev], dtype=np.int32)
    edge_f = np.array([e[1] for e in ev])
    edge_back = np.array([e[2] for e in ev], dtype=np.uint8)
    edge_pv = (3 + edge_f).astype(np.int32)
    E = edge_cam.size
    edge_off = np.concatenate([[0], np.cumsum([Bb.shape[0] if b else B.shape[0] for b in edge_back])]).astype(np.int64)
    obj = np.concatenate([(Bb if b else B).astype(np.float32) for b in edge_back], axis=0)
    img = np.empty((int(edge_off[-1]), 2), dtype=np.float32)
    nrng = _rng(seed, 51)
    for e in range(E):
        c, f = edge_cam[e], edge_f[e]
        X = Bb @ Rd.T + td if edge_back[e] else B
        Xc = (X @ frame_R[f].T + frame_t[f]) @ cR[c].T + ct[c]
        uv = project(cams["cam_model"][c], cams["cam_K5"][c], cams["cam_dist8"][c], cams["cam_xi"][c], Xc)
        img[edge_off[e]:edge_off[e + 1]] = (uv + noise_px * nrng.standard_normal(uv.shape)).astype(np.float32)
    p_true = np.concatenate([D_true[None], np.concatenate([log_so3_batch(frame_R), frame_t], axis=1)], axis=0)
    pert = _rng(seed, 52).standard_normal(p_true.shape)
    p_init = p_true.copy()
    p_init[:, :3] += init_rot * pert[:, :3]
    p_init[:, 3:] += init_trans * pert[:, 3:]
    p_init = p_init.astype(np.float32).astype(np.float64)
    return dict(n_cam=3, n_frame=n_frame, edge_cam=edge_cam, edge_pv=edge_pv, edge_off=edge_off, obj=obj, img=img,
                cam_model=cams["cam_model"], cam_K5=cams["cam_K5"], cam_dist8=cams["cam_dist8"], cam_ndist=cams["cam_ndist"],
                cam_xi=cams["cam_xi"], n_points=int(edge_off[-1]), edge_back=edge_back, cam_pose=cam_pose, cam_R=cR, cam_t=ct,
                ds_params_true=p_true.ravel(), ds_params_init=p_init.ravel())


def make_config(idx, n_frame=None, frame_stream=0):
    cfg = dict(CONFIGS[idx])
    models = cfg.pop("models", None)
    nC = cfg["n_cam"]
    if models == "omni":
        models = [OMNIDIRECTIONAL] * nC
    elif models == "mixed":
        models = [OMNIDIRECTIONAL if c % 2 else PINHOLE for c in range(nC)]
    if n_frame is not None:
        cfg["n_frame"] = n_frame
    return make_rig(models=models, frame_stream=frame_stream, **cfg)


def shard_rig(rig, rank, nranks):
    """Contiguous ranges of photo vertices, balanced by corner count.  Returns the local rig (photo vertices
    renumbered nC..nC+F_local-1, local edges in the original relative order) and the global vertex ids of the local
    frames.  Cameras (and their parameter slots) are replicated."""
    nC, F = rig["n_cam"], rig["n_frame"]
    n_per_edge = np.diff(rig["edge_off"])
    per_frame = np.bincount(rig["edge_pv"] - nC, weights=n_per_edge, minlength=F)
    cum = np.concatenate([[0], np.cumsum(per_frame)])
    total = cum[-1]
    bounds = [int(np.searchsorted(cum, total * r / nranks, side="left")) for r in range(nranks)] + [F]
    bounds[0] = 0
    f0, f1 = bounds[rank], bounds[rank + 1]
    sel = np.nonzero((rig["edge_pv"] - nC >= f0) & (rig["edge_pv"] - nC < f1))[0]
    off = rig["edge_off"]
    idx = np.concatenate([np.arange(off[e], off[e + 1]) for e in sel]) if sel.size else np.zeros(0, dtype=np.int64)
    local_off = np.concatenate([[0], np.cumsum(n_per_edge[sel])]).astype(np.int64)
    cam_part = slice(0, 6 * (nC - 1))
    fr_part = slice(6 * (nC - 1 + f0), 6 * (nC - 1 + f1))
    out = dict(rig)
    out.update(n_frame=f1 - f0, edge_cam=rig["edge_cam"][sel], edge_pv=(rig["edge_pv"][sel] - f0).astype(np.int32),
               edge_off=local_off, obj=rig["obj"][idx], img=rig["img"][idx],
               params_init=np.concatenate([rig["params_init"][cam_part], rig["params_init"][fr_part]]),
               params_true=np.concatenate([rig["params_true"][cam_part], rig["params_true"][fr_part]]),
               global_frame_range=(f0, f1), edge_index=sel, n_points=int(local_off[-1]))
    return out
