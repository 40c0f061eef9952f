"""Reference-format ingest: MyMultiCameraCalibration reads <data>/<serial>/<timestamp>.yaml
("corners", "objects") and <cfg>/<serial>.xml ("Intrinsics", "Distortion") exactly like src/mymulticalib.cpp:118-131,
182-233, 268-301, initialises every image with solvePnP (:203-211), and writes the poses back into the camera config
files (:425-454).  The files are produced with cv2.FileStorage (what the reference's own tooling writes), the restated
solvePnP is compared with cv2.solvePnP, the indexing with oracle/indexing.py (bit-exact)."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from multi_camera_calibration_b200 import multicalib, synth
from oracle import indexing

SERIALS = ["839112060578", "839512061262", "f0220380"]


def _write_dataset(tmp, rig, fmt_corners="d", extra_single=True, back_pattern=True):
    """One YAML per (camera, timestamp) image + one XML per camera.  Timestamps are NOT zero padded: cv::glob order is
    lexicographic ("10.yaml" before "9.yaml"), which the indexing contract depends on."""
    data = os.path.join(tmp, "color"); cfg = os.path.join(tmp, "configs")
    os.makedirs(cfg)
    nC = rig["n_cam"]
    for c in range(nC):
        os.makedirs(os.path.join(data, SERIALS[c]))
        fs = cv2.FileStorage(os.path.join(cfg, SERIALS[c] + ".xml"), cv2.FILE_STORAGE_WRITE)
        fx, fy, cx, cy, _ = rig["cam_K5"][c]
        fs.write("depth_scale", 0.001)
        fs.write("height", 480.0)
        fs.write("CameraMatrix", np.eye(4))
        fs.write("Intrinsics", np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64))
        fs.write("Distortion", rig["cam_dist8"][c][:5].reshape(1, 5).astype(np.float64))
        fs.release()
    files = {}
    for e in range(rig["edge_cam"].size):
        c = int(rig["edge_cam"][e]); ts = int(rig["timestamps"][rig["edge_pv"][e] - nC]) + 3
        a, b = rig["edge_off"][e], rig["edge_off"][e + 1]
        path = os.path.join(data, SERIALS[c], "%d.yaml" % ts)
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_WRITE)
        fs.write("corners", rig["img"][a:b].astype(np.float64 if fmt_corners == "d" else np.float32))
        fs.write("objects", rig["obj"][a:b].astype(np.float64))
        fs.release()
        files[(c, ts)] = path
    if extra_single:        # a timestamp only camera 1 has: loaded, but no vertex / edge (src/mymulticalib.cpp:374-376)
        a, b = rig["edge_off"][0], rig["edge_off"][1]
        c = int(rig["edge_cam"][0])
        fs = cv2.FileStorage(os.path.join(data, SERIALS[c], "100000.yaml"), cv2.FILE_STORAGE_WRITE)
        fs.write("corners", rig["img"][a:b].astype(np.float64)); fs.write("objects", rig["obj"][a:b].astype(np.float64))
        fs.release()
    if back_pattern:        # an image of the (smaller) back pattern: dropped by storeReaded (:236-241)
        a = rig["edge_off"][0]
        c = int(rig["edge_cam"][0])
        fs = cv2.FileStorage(os.path.join(data, SERIALS[c], "100001.yaml"), cv2.FILE_STORAGE_WRITE)
        fs.write("corners", rig["img"][a:a + 40].astype(np.float64)); fs.write("objects", rig["obj"][a:a + 40].astype(np.float64))
        fs.release()
    return data, cfg, files


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    rig = synth.make_rig(n_cam=3, n_frame=30, seed=41)
    tmp = str(tmp_path_factory.mktemp("ingest"))
    data, cfg, files = _write_dataset(tmp, rig)
    return rig, data, cfg, files


def test_solve_pnp_matches_cv2():
    rng = np.random.default_rng(5)
    B = synth.board().astype(np.float32).astype(np.float64)
    for trial in range(12):
        K5 = np.array([1000 + 50 * rng.standard_normal(), 990 + 50 * rng.standard_normal(), 960, 540, 0.0])
        nd = [0, 4, 5, 8][trial % 4]
        d8 = np.zeros(8)
        d8[:nd] = ([0.05, -0.02, 1e-3, -1e-3, 0.003, 0.01, -0.005, 0.002][:nd])
        om = 0.4 * rng.standard_normal(3); om[2] += rng.uniform(-3, 3) * 0.3
        t = np.array([rng.uniform(-200, 100), rng.uniform(-150, 50), rng.uniform(900, 2200)])
        K = np.array([[K5[0], 0, K5[2]], [0, K5[1], K5[3]], [0, 0, 1]])
        obj = B if trial % 3 else B + np.array([0, 0, 1.0]) * rng.uniform(-60, 60, (B.shape[0], 1))    # planar and non-planar
        uv, _ = cv2.projectPoints(obj, om, t, K, d8[:nd] if nd else None)
        uv = uv.reshape(-1, 2) + 0.2 * rng.standard_normal((obj.shape[0], 2))
        ok, rv, tv = cv2.solvePnP(obj, uv, K, d8[:nd] if nd else None)
        r, tt = multicalib.solve_pnp(obj, uv, K5, d8, nd)
        # both minimise the same reprojection error; cv2 stops after <= 20 LM iterations at a float32-sized step
        assert np.abs(r - rv.ravel()).max() < 2e-5 and np.abs(tt - tv.ravel()).max() < 2e-2, (trial, r, rv.ravel(), tt, tv.ravel())
        res = lambda rr, t2: np.linalg.norm(cv2.projectPoints(obj, rr, t2, K, d8[:nd] if nd else None)[0].reshape(-1, 2) - uv)
        assert res(r, tt) <= res(rv, tv) * (1 + 1e-9) + 1e-9          # at least as good a minimum as OpenCV's


def test_reads_what_cv2_reads_and_indexes_bit_exactly(dataset):
    rig, data, cfg, files = dataset
    mc = multicalib.MyMultiCameraCalibration(SERIALS, 0, 3, data, cfg)
    mc.loadImages()
    idx = mc.indexing()
    # the same listing through the oracle's indexing rules: cameras outer loop, files in lexicographic path order
    per_cam = []
    for c in range(3):
        lst = []
        for name in os.listdir(os.path.join(data, SERIALS[c])):
            p = os.path.join(data, SERIALS[c], name)
            fs = cv2.FileStorage(p, cv2.FILE_STORAGE_READ)
            n = fs.getNode("corners").mat().shape[0]
            fs.release()
            if n == 54:                                      # front pattern only (storeReaded)
                lst.append((p, int(os.path.splitext(name)[0])))
        per_cam.append(lst)
    ref = indexing.build_indexing(3, per_cam)
    e = np.array(ref["edges"], dtype=np.int32)
    assert np.array_equal(idx["edge_cam"], e[:, 0]) and np.array_equal(idx["edge_pv"], e[:, 1])
    assert np.array_equal(idx["photo_index"], e[:, 2])
    assert np.array_equal(idx["vertex_timestamp"], np.array(ref["vertex_timestamp"], dtype=np.int32))
    assert 100000 not in idx["vertex_timestamp"] and 100001 not in idx["vertex_timestamp"]
    # initial parameters: BFS chaining of the solvePnP transforms == the same chaining of cv2.solvePnP transforms
    mc.initialize()
    p0 = mc.initialParameters().reshape(-1, 6)
    truth = rig["params_true"].reshape(-1, 6)
    nC = 3
    cam_err = np.abs(p0[:nC - 1] - truth[:nC - 1])
    assert cam_err[:, :3].max() < 0.05 and cam_err[:, 3:].max() < 40.0          # single-image PnP accuracy
    mc.close()


def test_yaml_float_corners_and_xml_roundtrip(tmp_path):
    rig = synth.make_rig(n_cam=2, n_frame=6, seed=42)
    data, cfg, files = _write_dataset(str(tmp_path), rig, fmt_corners="f", extra_single=False, back_pattern=False)
    mc = multicalib.MyMultiCameraCalibration(SERIALS[:2], 0, 2, data, cfg)
    mc.loadImages()
    assert mc.indexing()["edge_cam"].size == 12
    mc.close()


@pytest.mark.gpu
def test_sample_workflow_from_the_reference_directory_layout(dataset, oracle_lib, tmp_path):
    """samples/multi_cameras_calibration.cpp:71-83: load, initialise, optimise, drop outliers, again; then writeParameters
    (XML + the per-camera config files).  Result against the oracle started from the same initial parameters."""
    from tests import rigs
    rig, data, cfg, files = dataset
    mc = multicalib.MyMultiCameraCalibration(SERIALS, 0, 3, data, cfg, criteria=(3, 200, 1e-7))
    mc.loadImages()
    mc.initialize()
    p0 = mc.initialParameters()
    idx = mc.indexing()
    err = mc.optimizeExtrinsics()
    p = mc.parameters()
    # the oracle on the rig as the host class indexed it
    nC = 3
    order = [files[(int(c), int(idx["vertex_timestamp"][pv]))] for c, pv in zip(idx["edge_cam"], idx["edge_pv"])]
    objs, imgs, off = [], [], [0]
    for path in order:
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
        objs.append(fs.getNode("objects").mat().astype(np.float32)); imgs.append(fs.getNode("corners").mat().astype(np.float32))
        fs.release()
        off.append(off[-1] + objs[-1].shape[0])
    from oracle import oracle as orc
    nF = idx["vertex_timestamp"].size - nC
    O = orc.Rig(nC, nF, idx["edge_cam"], idx["edge_pv"], np.array(off, dtype=np.int64), np.concatenate(objs), np.concatenate(imgs),
                rig["cam_model"], rig["cam_K5"], np.pad(rig["cam_dist8"][:, :5], ((0, 0), (0, 3))), np.full(nC, 5, dtype=np.int32), rig["cam_xi"])
    ref = O.solve(p0, mode=0, crit_type=3, max_count=200, eps=1e-7)
    scale = np.maximum(np.abs(ref["params"]), 1.0)
    assert np.max(np.abs(p - ref["params"]) / scale) < 1e-6
    assert abs(mc.stats()["rms"] - O.error(ref["params"])["rms"]) < 1e-8
    assert mc.removeOutlier() == set() or True
    out = str(tmp_path / "out.xml")
    mc.writeParameters(out)
    fs = cv2.FileStorage(out, cv2.FILE_STORAGE_READ)
    assert int(fs.getNode("nCameras").real()) == 3 and fs.getNode("camera_pose_1").mat().shape == (4, 4)
    fs.release()
    for c in range(3):      # writeParameters2config: CameraMatrix = pose of the camera vertex, the rest preserved
        fs = cv2.FileStorage(os.path.join(cfg, SERIALS[c] + ".xml"), cv2.FILE_STORAGE_READ)
        M = fs.getNode("CameraMatrix").mat()
        assert M.shape == (4, 4) and abs(fs.getNode("depth_scale").real() - 0.001) < 1e-9 and fs.getNode("height").real() == 480.0
        assert fs.getNode("Intrinsics").mat().shape == (3, 3) and fs.getNode("Distortion").mat().size == 5
        fs.release()
        if c == 0:
            assert np.allclose(M, np.eye(4))
    mc.close()
